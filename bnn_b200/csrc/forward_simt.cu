// Kernel (a)+(e), FP32 SIMT variant: fused S x N batched small-MLP forward with
// the predictive Welford reduction in the epilogue.
//
// Replaces the reference's `for i in range(S): pred[i] = nns[i](X)` loops
// (BNN_SGDMC.py:131-137, BNN_BBB.py:143-149, BNN_Dropout.py:78-84,
// BNN_CDropout.py:158-164, BNN_SVI.py:70-76; inner op util.py:31-33) and
// `preds.mean(0), preds.var(0)` (BNN.py:36-37).
//
// Work decomposition: one CTA = (tile of TN points) x (group of posterior samples).
// The CTA walks its samples; per sample it runs the whole MLP for its TN points with
// the activations resident in shared memory (h[k][n], ping-pong) and never writes
// them to HBM.  Weights are streamed in K-chunks through a 3-stage shared-memory ring
// by a dedicated producer warp using the bulk-copy engine (cp.async.bulk + mbarrier);
// 8 consumer warps do register-tiled FP32 FFMA (thread tile (4*NI/4.. rows) x RN points,
// lane grid 4 x 8).  The last layer's outputs feed per-point FP64 Welford state kept in
// shared memory; only (mean, M2) per point leave the CTA (and pred[S,N,O] if asked).
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint

#include "common.cuh"
#include "philox.cuh"

namespace bnn {

constexpr int kComputeWarps = 8;
constexpr int kKC = 32;  // K-chunk: 32 floats = one 128-byte swizzle row
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kThreads = kComputeThreads + 32;  // + producer warp
constexpr int kMaxStages = 3;
constexpr int kSmemBudget = 227 * 1024;

struct FwdLayer {
  int in, out, ld;
  int ni;        // thread tile rows = ni (tile height 4*ni)
  int tiles;     // ceil(out / (4*ni))
  int rounds;    // ceil(tiles / kComputeWarps)
  int chunks;    // ceil(in / KC)
  int mask_off;  // offset of this layer's INPUT mask, -1 = none
  int box_rows;  // rows of one staged block (TMA box height)
  long long off;
};

struct FwdParams {
  int n_lin, act, nstages, stage_floats, nout, d4, hrows;
  FwdLayer L[BNN_MAX_LINEAR];
  alignas(64) CUtensorMap tmap[BNN_MAX_LINEAR];  // per layer: dims (in, out, S), box (32, stage rows, 1), 128B swizzle
  const float* X;
  long long N;
  const float* W;
  long long w_stride;
  int S, G, Sg;
  int mask_mode, mask_len;
  const float* mask;
  long long mask_stride;
  float p_drop[BNN_MAX_LINEAR];
  float temperature;
  unsigned long long seed;
  long long sample_offset;
  float* pred;
  double* part;  // [G][2][M]
  int sm_msk, sm_wf, sm_xs, sm_h0, sm_h1, sm_ring;
};

// acc[i][j] += W[row_i][k] * h[k][n_j] over one staged K-chunk
// The staged block is dense [rows][32 floats] written by TMA with the 128-byte swizzle: the 16-byte
// granule g of row r sits at granule g ^ (r & 7), which makes the 4 rows a warp reads per LDS.128
// (r, r+1, r+2, r+3) land in distinct banks.  All addresses are 32-bit shared-space byte addresses.
template <int NI, int RN>
__device__ __forceinline__ void tile_fma(float (&acc)[NI][RN], uint32_t wst, uint32_t h, int kq_count, int row0, int ln) {
  constexpr int TN = 8 * RN;
  const uint32_t wrow = wst + (uint32_t)row0 * (kKC * 4);  // row0 = first staged row of this thread (stage_row + lo)
  const uint32_t swz0 = (uint32_t)(row0 & 7);
  const uint32_t hcol = h + (RN == 2 ? ln * 8 : ln * 16);
#pragma unroll 2
  for (int kq = 0; kq < kq_count; ++kq) {
    float4 w[NI];
    const uint32_t g0 = (((uint32_t)kq ^ swz0) << 4), g1 = (((uint32_t)kq ^ swz0 ^ 4u) << 4);  // rows r and r+4
#pragma unroll
    for (int i = 0; i < NI; ++i) w[i] = lds128(wrow + i * 4 * (kKC * 4) + ((i & 1) ? g1 : g0));
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float a[RN];
      const uint32_t hp = hcol + (uint32_t)(kq * 4 + kk) * (TN * 4);
      if constexpr (RN == 8) {
        const float4 a0 = lds128(hp);
        const float4 a1 = lds128(hp + 128);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      } else if constexpr (RN == 4) {
        const float4 a0 = lds128(hp);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      } else {
        const float2 a0 = lds64(hp);
        a[0] = a0.x; a[1] = a0.y;
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const float wv = kk == 0 ? w[i].x : (kk == 1 ? w[i].y : (kk == 2 ? w[i].z : w[i].w));
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(wv, a[j], acc[i][j]);
      }
    }
  }
}

struct PipeState {
  int stage;
  uint32_t phase;
  __device__ __forceinline__ void advance(int nstages) {
    if (++stage == nstages) { stage = 0; phase ^= 1u; }
  }
};

// one TMA tile load: box of the layer's tensor map at (k0, row0, sample) -> smem, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// one (layer, round): this warp's tile over all K-chunks, then bias/act/mask epilogue into hout
template <int NI, int RN>
__device__ __forceinline__ void layer_round(const FwdParams& p, const FwdLayer& L, int l, int round, const float* __restrict__ Ws,
                                            const float* __restrict__ hin, float* __restrict__ hout, const float* __restrict__ msk,
                                            const float* ring, uint64_t* full, uint64_t* empty, PipeState& ps, int warp, int lane) {
  constexpr int TN = 8 * RN;
  const int lo = lane >> 3, ln = lane & 7;
  const int tile = round * kComputeWarps + warp;
  const bool valid = tile < L.tiles;
  const int row_base = tile * 4 * NI;                        // first row of this warp's tile
  const int stage_row = (tile - round * kComputeWarps) * 4 * NI;  // row inside the staged block

  float acc[NI][RN];
  float bias[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int row = row_base + i * 4 + lo;
    bias[i] = (valid && row < L.out) ? __ldg(Ws + L.off + (long long)row * L.ld + L.in) : 0.0f;
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.0f;
  }

  for (int c = 0; c < L.chunks; ++c) {
    const int k0 = c * kKC;
    const int kw = min(kKC, roundup4(L.in - k0));
    mbar_wait(&full[ps.stage], ps.phase);
    if (valid)
      tile_fma<NI, RN>(acc, smem_u32(ring) + (uint32_t)ps.stage * (uint32_t)(p.stage_floats * 4),
                       smem_u32(hin) + (uint32_t)k0 * (TN * 4), kw >> 2, stage_row + lo, ln);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[ps.stage]);
    ps.advance(p.nstages);
  }

  if (valid) {
    const bool last = (l == p.n_lin - 1);
    const int next_mask = last ? -1 : p.L[l + 1].mask_off;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int row = row_base + i * 4 + lo;
      if (row < L.out) {
        const float m = next_mask >= 0 ? msk[next_mask + row] : 1.0f;
        float v[RN];
#pragma unroll
        for (int j = 0; j < RN; ++j) {
          float t = acc[i][j] + bias[i];
          if (!last) t = apply_act(t, p.act);
          v[j] = next_mask >= 0 ? t * m : t;
        }
        const uint32_t dst = smem_u32(hout) + (uint32_t)row * (TN * 4);
        if constexpr (RN == 8) {
          sts128(dst + ln * 16, make_float4(v[0], v[1], v[2], v[3]));
          sts128(dst + 128 + ln * 16, make_float4(v[4], v[5], v[6], v[7]));
        } else if constexpr (RN == 4) {
          sts128(dst + ln * 16, make_float4(v[0], v[1], v[2], v[3]));
        } else {
          sts64(dst + ln * 8, make_float2(v[0], v[1]));
        }
      }
    }
  }
}

__device__ __forceinline__ float mask_value(const FwdParams& p, long long s_global, int s_local, int layer, int j_global) {
  if (p.mask_mode == BNN_MASK_EXTERNAL) return __ldg(p.mask + (long long)s_local * p.mask_stride + j_global);
  const uint32_t tag = p.mask_mode == BNN_MASK_BERNOULLI ? STREAM_BERNOULLI : STREAM_CONCRETE;
  const u32x4 blk = philox_block(p.seed, (uint32_t)s_global, tag, (uint64_t)(j_global >> 2));
  const float u = u24(pick_word(blk, j_global & 3));
  return p.mask_mode == BNN_MASK_BERNOULLI ? bernoulli_keep(u, p.p_drop[layer]) : concrete_keep(u, p.p_drop[layer], p.temperature);
}

template <int RN>
__global__ void __launch_bounds__(kThreads, 1) fwd_simt_kernel(const __grid_constant__ FwdParams p) {
  constexpr int TN = 8 * RN;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  float* msk = reinterpret_cast<float*>(smem + p.sm_msk);
  double* wf_mean = reinterpret_cast<double*>(smem + p.sm_wf);
  double* wf_m2 = wf_mean + p.nout * TN;
  float* xs = reinterpret_cast<float*>(smem + p.sm_xs);
  float* hbuf[2] = {reinterpret_cast<float*>(smem + p.sm_h0), reinterpret_cast<float*>(smem + p.sm_h1)};
  float* ring = reinterpret_cast<float*>(smem + p.sm_ring);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n0 = (long long)blockIdx.x * TN;
  const int g = blockIdx.y;
  const int s0 = g * p.Sg, s1 = min(p.S, s0 + p.Sg);

  if (tid == 0) {
    for (int i = 0; i < p.nstages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kComputeWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == kComputeWarps) {
    // ===================== producer warp: weights -> smem ring (one TMA per stage) =====================
    if (lane == 0) {
      PipeState ps{0, 0};
      const bool shared_w = (p.w_stride == 0);
      for (int s = s0; s < s1; ++s) {
        for (int l = 0; l < p.n_lin; ++l) {
          const FwdLayer& L = p.L[l];
          const uint32_t stage_bytes = (uint32_t)(L.box_rows * kKC * 4);
          for (int r = 0; r < L.rounds; ++r) {
            const int row0 = r * L.box_rows;
            for (int c = 0; c < L.chunks; ++c) {
              mbar_wait_backoff(&empty[ps.stage], ps.phase ^ 1u);
              mbar_arrive_expect_tx(&full[ps.stage], stage_bytes);  // OOB rows/cols are zero-filled and counted
              tma_load_3d(ring + (size_t)ps.stage * p.stage_floats, &p.tmap[l], c * kKC, row0, shared_w ? 0 : s,
                          &full[ps.stage]);
              ps.advance(p.nstages);
            }
          }
        }
      }
    }
    return;
  }

  // ========================= consumer warps ===================================
  PipeState ps{0, 0};
  const int D = p.L[0].in;
  // X tile, transposed to xs[d][n]; rows beyond N and pad rows are zero
  for (int idx = tid; idx < p.d4 * TN; idx += kComputeThreads) xs[idx] = 0.0f;
  named_bar_sync(1, kComputeThreads);
  {
    const long long remaining = p.N - n0;
    const int npts = remaining < TN ? (int)remaining : TN;
    const float* xg = p.X + n0 * D;
    for (int idx = tid; idx < npts * D; idx += kComputeThreads) {
      const int n = idx / D, d = idx - n * D;
      xs[d * TN + n] = __ldg(xg + idx);
    }
  }
  const int nstate = p.nout * TN;
  for (int idx = tid; idx < nstate; idx += kComputeThreads) {
    wf_mean[idx] = 0.0;
    wf_m2[idx] = 0.0;
  }
  named_bar_sync(1, kComputeThreads);

  for (int s = s0; s < s1; ++s) {
    const float* Ws = p.W + (long long)s * p.w_stride;
    const long long s_global = p.sample_offset + s;
    if (p.mask_mode != BNN_MASK_NONE) {
      for (int l = 0; l < p.n_lin; ++l) {
        const int mo = p.L[l].mask_off;
        if (mo < 0) continue;
        for (int j = tid; j < p.L[l].in; j += kComputeThreads) msk[mo + j] = mask_value(p, s_global, s, l, mo + j);
      }
      named_bar_sync(1, kComputeThreads);
    }
    const float* hin = xs;
    int cur = 0;
    if (p.L[0].mask_off >= 0) {
      float* h = hbuf[0];
      for (int idx = tid; idx < p.d4 * TN; idx += kComputeThreads) {
        const int d = idx / TN;
        h[idx] = d < D ? xs[idx] * msk[p.L[0].mask_off + d] : 0.0f;
      }
      named_bar_sync(1, kComputeThreads);
      hin = hbuf[0];
      cur = 1;
    }
    for (int l = 0; l < p.n_lin; ++l) {
      const FwdLayer& L = p.L[l];
      float* hout = hbuf[cur];
      for (int r = 0; r < L.rounds; ++r) {
        switch (L.ni) {
          case 8: layer_round<8, RN>(p, L, l, r, Ws, hin, hout, msk, ring, full, empty, ps, warp, lane); break;
          case 4: layer_round<4, RN>(p, L, l, r, Ws, hin, hout, msk, ring, full, empty, ps, warp, lane); break;
          case 2: layer_round<2, RN>(p, L, l, r, Ws, hin, hout, msk, ring, full, empty, ps, warp, lane); break;
          default: layer_round<1, RN>(p, L, l, r, Ws, hin, hout, msk, ring, full, empty, ps, warp, lane); break;
        }
      }
      // rows [out, roundup4(out)) feed the next layer's 4-wide K loop: keep them zero
      const int pad = roundup4(L.out) - L.out;
      for (int idx = tid; idx < pad * TN; idx += kComputeThreads) hout[(size_t)L.out * TN + idx] = 0.0f;
      named_bar_sync(1, kComputeThreads);
      hin = hout;
      cur ^= 1;
    }
    // hin now holds y[o][n]: Welford update (+ optional pred store)
    const double inv_cnt = 1.0 / (double)(s - s0 + 1);
    for (int idx = tid; idx < nstate; idx += kComputeThreads) {
      const int o = idx / TN, n = idx - o * TN;
      const float y = hin[idx];
      const double d0 = (double)y - wf_mean[idx];
      const double mu = wf_mean[idx] + d0 * inv_cnt;
      wf_mean[idx] = mu;
      wf_m2[idx] += d0 * ((double)y - mu);
      if (p.pred != nullptr && n0 + n < p.N) p.pred[((long long)s * p.N + n0 + n) * p.nout + o] = y;
    }
    named_bar_sync(1, kComputeThreads);
  }

  if (p.part != nullptr) {
    const long long M = p.N * p.nout;
    double* pm = p.part + (long long)g * 2 * M;
    for (int idx = tid; idx < nstate; idx += kComputeThreads) {
      const int o = idx / TN, n = idx - o * TN;
      if (n0 + n < p.N) {
        const long long m = (n0 + n) * p.nout + o;
        pm[m] = wf_mean[idx];
        pm[M + m] = wf_m2[idx];
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------
// Thread-tile height.  Tall tiles (NI = 8: 8 rows x RN points per thread) are what keeps the loop FMA-bound:
// per 4-wide K step a warp issues NI*32 FFMA against NI + 8 LDS.128, so NI = 1 is shared-memory-bound by ~4x.
// Only shrink the tile when the layer is too narrow to give every compute warp a tile.
static int choose_ni(int out) {
  const int cand[4] = {8, 4, 2, 1};
  for (int c = 0; c < 4; ++c) {
    const int ni = cand[c];
    const int tiles = (out + 4 * ni - 1) / (4 * ni);
    if (tiles >= kComputeWarps - 1 || ni == 1) return ni;  // (almost) every warp busy
  }
  return 1;
}

struct FwdPlan {
  FwdParams p;
  int RN;
  size_t smem;
  long long tiles;
};

static int plan_forward(const BnnForwardArgs* a, const Layout& L, FwdPlan* plan) {
  FwdParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.n_lin = L.n_lin;
  p.act = L.act;
  p.nout = L.out[L.n_lin - 1];
  p.d4 = roundup4(L.in[0]);
  int hrows = 4;
  int stage_rows = 4;
  int mask_off[BNN_MAX_LINEAR];
  p.mask_mode = a->mask.mode;
  p.mask_len = a->mask.mode == BNN_MASK_NONE ? 0 : make_mask_offsets(L, a->mask.layer_masked, mask_off);
  for (int l = 0; l < L.n_lin; ++l) {
    FwdLayer& f = p.L[l];
    f.in = L.in[l]; f.out = L.out[l]; f.ld = L.ld[l]; f.off = L.off[l];
    f.ni = choose_ni(f.out);
    f.tiles = (f.out + 4 * f.ni - 1) / (4 * f.ni);
    f.rounds = (f.tiles + kComputeWarps - 1) / kComputeWarps;
    f.mask_off = a->mask.mode == BNN_MASK_NONE ? -1 : mask_off[l];
    if (roundup4(f.out) > hrows) hrows = roundup4(f.out);
    // rows a consumer may touch in one staged block: whole tiles, even if the last one is ragged
    const int rr = (f.tiles < kComputeWarps ? f.tiles : kComputeWarps) * 4 * f.ni;
    f.box_rows = rr;
    if (rr > stage_rows) stage_rows = rr;
    p.p_drop[l] = a->mask.p_drop[l];
  }
  if (p.d4 > hrows && p.L[0].mask_off >= 0) hrows = p.d4;  // masked input is staged in h0
  p.hrows = hrows;
  if (p.nout > BNN_MAX_FUSED_NOUT) BNN_FAIL("nout=%d exceeds BNN_MAX_FUSED_NOUT=%d", p.nout, BNN_MAX_FUSED_NOUT);

  // pick the largest point tile (and ring depth) that fits the 227 KB shared-memory budget
  stage_rows = (stage_rows + 7) & ~7;  // whole 1024-byte swizzle atoms
  p.stage_floats = stage_rows * kKC;
  const int tn_cand[3] = {64, 32, 16};
  bool ok = false;
  for (int ti = 0; ti < 3 && !ok; ++ti) {
    for (int nst = kMaxStages; nst >= 2 && !ok; --nst) {
      const int TN = tn_cand[ti];
      size_t o = 128;  // barriers
      p.sm_msk = (int)o; o += (size_t)((p.mask_len + 3) & ~3) * 4;
      o = (o + 15) & ~(size_t)15;
      p.sm_wf = (int)o; o += (size_t)p.nout * TN * 16;
      p.sm_xs = (int)o; o += (size_t)p.d4 * TN * 4;
      p.sm_h0 = (int)o; o += (size_t)hrows * TN * 4;
      p.sm_h1 = (int)o; o += (size_t)hrows * TN * 4;
      o = (o + 1023) & ~(size_t)1023;  // 128B-swizzled TMA destinations need 1024-byte alignment
      p.sm_ring = (int)o;
      o += (size_t)nst * p.stage_floats * 4;
      if (o <= (size_t)kSmemBudget) {
        ok = true;
        plan->RN = TN / 8;
        plan->smem = o;
        p.nstages = nst;
      }
    }
  }
  if (!ok) BNN_FAIL("network too wide for the SIMT forward (shared-memory budget exceeded)");
  for (int l = 0; l < L.n_lin; ++l) p.L[l].chunks = (p.L[l].in + kKC - 1) / kKC;

  const int TN = plan->RN * 8;
  plan->tiles = (a->N + TN - 1) / TN;
  // sample groups: enough CTAs to fill the machine ~2x when the point tiles alone cannot
  const long long target = 2LL * sm_count();
  long long G = 1;
  if (plan->tiles < target) G = (target + plan->tiles - 1) / plan->tiles;
  if (G > a->S) G = a->S;
  if (G > 65535) G = 65535;
  const long long Sg = (a->S + G - 1) / G;
  G = (a->S + Sg - 1) / Sg;
  p.G = (int)G;
  p.Sg = (int)Sg;
  p.S = (int)a->S;
  p.X = a->X; p.N = a->N; p.W = a->W; p.w_stride = a->w_stride;
  p.mask = a->mask.mask; p.mask_stride = a->mask.mask_stride;
  p.temperature = a->mask.temperature; p.seed = a->mask.seed; p.sample_offset = a->mask.sample_offset;
  p.pred = a->pred;
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// per layer: 3-D view (k, row, sample) of the bank; K beyond `in` (bias + padding columns) and rows beyond
// `out` are out of bounds for the map, so TMA zero-fills them and the FMA loop needs no tail handling
static int make_tensor_maps(const BnnForwardArgs* a, FwdParams& p) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) BNN_FAIL("cuTensorMapEncodeTiled is not available from this driver");
  for (int l = 0; l < p.n_lin; ++l) {
    const FwdLayer& L = p.L[l];
    const bool shared_w = a->w_stride == 0;
    cuuint64_t gdim[3] = {(cuuint64_t)L.in, (cuuint64_t)L.out, (cuuint64_t)(shared_w ? 1 : a->S)};
    cuuint64_t gstr[2] = {(cuuint64_t)L.ld * 4, (cuuint64_t)(shared_w ? (long long)L.out * L.ld : a->w_stride) * 4};
    cuuint32_t box[3] = {(cuuint32_t)kKC, (cuuint32_t)L.box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    void* base = const_cast<float*>(a->W) + L.off;
    const CUresult r = enc(&p.tmap[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) BNN_FAIL("cuTensorMapEncodeTiled failed for layer %d (CUresult %d)", l, (int)r);
  }
  return 0;
}

long long forward_simt_workspace_bytes(const BnnForwardArgs* a) {
  Layout L;
  if (make_layout(&a->desc, &L)) return -1;
  FwdPlan plan;
  if (plan_forward(a, L, &plan)) return -1;
  return (long long)plan.p.G * 2 * a->N * plan.p.nout * (long long)sizeof(double);
}

int finalize_moments(const double* part, const int* counts_dev_or_null, int G, int Sg, long long S, long long M, float* mean,
                     float* var, double* part_mean, double* part_m2, cudaStream_t st);

int forward_simt(const BnnForwardArgs* a, cudaStream_t st) {
  Layout L;
  if (make_layout(&a->desc, &L)) return -1;
  FwdPlan plan;
  if (plan_forward(a, L, &plan)) return -1;
  FwdParams& p = plan.p;
  const bool want_moments = a->mean || a->var || a->part_mean || a->part_m2;
  const long long M = a->N * p.nout;
  if (want_moments) {
    const long long need = (long long)p.G * 2 * M * (long long)sizeof(double);
    if (!a->workspace || a->workspace_bytes < need)
      BNN_FAIL("workspace too small: need %lld bytes, got %lld", need, (long long)a->workspace_bytes);
    p.part = reinterpret_cast<double*>(a->workspace);
  } else {
    p.part = nullptr;
  }
  if (make_tensor_maps(a, p)) return -1;
  dim3 grid((unsigned)plan.tiles, (unsigned)p.G);
  auto launch = [&](auto kern) -> int {
    BNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    kern<<<grid, kThreads, plan.smem, st>>>(p);
    BNN_CUDA(cudaGetLastError());
    return 0;
  };
  int rc = 0;
  if (plan.RN == 8) rc = launch(fwd_simt_kernel<8>);
  else if (plan.RN == 4) rc = launch(fwd_simt_kernel<4>);
  else rc = launch(fwd_simt_kernel<2>);
  if (rc) return rc;
  if (want_moments)
    return finalize_moments(p.part, nullptr, p.G, p.Sg, a->S, M, a->mean, a->var, a->part_mean, a->part_m2, st);
  return 0;
}

}  // namespace bnn
