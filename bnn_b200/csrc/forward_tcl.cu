// Kernel (a)+(e), tensor-core variant for DEEP / WIDE networks (cfg 5's 90-512-512-512-1 and anything outside the
// [D <= 64, H <= 256, two hidden layers] family of forward_tc.cu): any number of hidden layers, widths up to 1024.
// Replaces the same reference loops as forward_simt.cu (BNN_SGDMC.py:131-137 et al. + BNN.py:36-37).
//
// At these widths one network's operand images (8 bytes per weight) are megabytes, so nothing is resident: a CTA pair
// (cta_group::2, M = 256 points) walks the layers of ONE (sample, 256-point tile) item and STREAMS both operands through a
// 3-stage shared-memory ring in K-panels of 32:
//   A = the tile's activations of the previous layer, B = the sample's weights, both pre-split into the three images of the
//   FP32-grade scheme (tf32(x) | bf16(x) | bf16(x - tf32(x)):  a b ~= a_h b_h [kind::tf32] + a_l b + a b_l [kind::f16]),
//   stored K-major in 16-byte granules so that one cp.async.bulk per image brings a panel in.
// Per layer the pair computes N-tiles of up to 256 output features into one of TWO accumulators in tensor memory (2 x 256
// columns): the 8 epilogue warps drain tile j (bias, activation, split into the three images of the NEXT layer's A operand,
// written to a per-CTA scratch in global memory that stays in L2) while the MMAs of tile j + 1 run.  The last Linear is a dot
// product in the epilogue of the last hidden layer; outputs feed the same FP64 Welford state as the other forwards.
// Each CTA only ever reads activations it wrote itself, so there is no inter-CTA dependency: one launch does all layers.
//
// Roofline note: the operands move 16 bytes per point-feature pair and panel (8 for A, 8 for B) against 2 * 4 = 8
// bf16-equivalent MMA units, i.e. 64 B/clk/SM when the tensor pipe is busy -- the kernel is bound by the L2 -> SM path, not
// by the tensor pipe.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace bnn {

constexpr int kTlThreads = 320;       // warp 0 loader, warp 1 MMA issuer (leader) / forwarder (peer), warps 2-9 epilogue
constexpr int kTlKP = 32;             // K panel
constexpr int kTlStages = 3;
constexpr int kTlStageBytes = 64 * 1024;   // A: tf32 16 KB | bf16 8 KB | bf16-lo 8 KB ; B: the same for <= 128 rows
constexpr int kTlMaxNT = 4;           // N-tiles of 256 per layer (width <= 1024)

struct TlLayer {
  int in, out, Kp, N1, n_nt, n_kp;
  long long w_off;                    // floats from the sample's image to this layer's first N-tile
  long long b_off, b_ld;              // bias in the BANK: W[s * w_stride + b_off + row * b_ld]
};

struct TlParams {
  int Lh, act, O, D;
  TlLayer ly[BNN_MAX_LINEAR];
  long long last_off, last_ld;        // last Linear in the bank: row o at last_off + o * last_ld (bias at column H)
  int H;                              // width feeding the last Linear
  const float* ximg;                  // [n_tiles128][2 * 128 * Kp0]
  const float* wimg;                  // [S][wimg_stride]
  long long wimg_stride;
  const float* W;                     // bank (biases, last Linear)
  long long w_stride;
  float* scratch;                     // [CTAs][2][2 * 128 * Kp_max]
  long long scratch_stride;           // floats per buffer
  long long N;
  int S, G, Sg, tiles_per_slot, n_tiles;   // pair tiles of 256 points
  float* pred;
  double* part;                       // [G][2][N * O]
};

// ---- pair-mode helpers (same instructions as forward_tc.cu) -----------------------------------------------------------------
__device__ __forceinline__ uint32_t tl_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tl_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tl_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tl_mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tl_mma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tl_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int kAct>
__device__ __forceinline__ float tl_act(float v) {
  if constexpr (kAct == BNN_ACT_RELU) return fmaxf(v, 0.0f);
  else if constexpr (kAct == BNN_ACT_TANH) {   // 1 - 2 / (exp(2 v) + 1) through ex2 / rcp: |err| <= 2e-7 (see forward_tc.cu)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
  } else if constexpr (kAct == BNN_ACT_SIGMOID) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return r;
  } else return v;
}

// ---- operand image builders ------------------------------------------------------------------------------------------------
// X [N, D] -> per 128-row tile the three images of a [128, Kp0] A operand (zero beyond N / D)
__global__ void __launch_bounds__(256) tl_prep_x_kernel(const float* __restrict__ X, long long N, int D, int Kp, long long n_tiles128,
                                                        float* __restrict__ img) {
  const long long quads = n_tiles128 * 128 * (Kp / 4);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(q % 128);
    const long long t = q / 128;
    const int kc = (int)(t % (Kp / 4));
    const long long tile = t / (Kp / 4);
    const long long n = tile * 128 + row;
    float x[4], hi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = kc * 4 + e;
      x[e] = (n < N && k < D) ? __ldg(X + n * D + k) : 0.0f;
      hi[e] = tc::tf32_hi(x[e]);
    }
    float* base = img + tile * (2LL * 128 * Kp);
    *reinterpret_cast<float4*>(base + ((long long)kc * 128 + row) * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    // bf16 images: granule = 8 k; this quad fills half of one (words (kc & 1) * 2, +1)
    uint32_t* bf = reinterpret_cast<uint32_t*>(base + 128LL * Kp) + ((long long)(kc >> 1) * 128 + row) * 4 + (kc & 1) * 2;
    uint32_t* bl = bf + 64LL * Kp;    // 128 * Kp / 2 words further
    bf[0] = tc::pack_bf16x2(x[0], x[1]); bf[1] = tc::pack_bf16x2(x[2], x[3]);
    bl[0] = tc::pack_bf16x2(x[0] - hi[0], x[1] - hi[1]); bl[1] = tc::pack_bf16x2(x[2] - hi[2], x[3] - hi[3]);
  }
}

// one sample's hidden-layer weights -> per layer, per N-tile of <= 256 rows, per CTA half: three K-major images [kc][row][4]
struct TlPrepW {
  int Lh;
  TlLayer ly[BNN_MAX_LINEAR];
  long long l_off[BNN_MAX_LINEAR];    // bank offset of each Linear
  int l_ld[BNN_MAX_LINEAR];
  long long w_stride, wimg_stride;
};
__global__ void __launch_bounds__(256) tl_prep_w_kernel(const __grid_constant__ TlPrepW q, const float* __restrict__ W, float* __restrict__ img) {
  const long long s = blockIdx.y;
  const float* Ws = W + s * q.w_stride;
  float* out = img + s * q.wimg_stride;
  for (int l = 0; l < q.Lh; ++l) {
    const TlLayer& Y = q.ly[l];
    const long long quads = (long long)Y.N1 * (Y.Kp / 4);       // one quad = 4 consecutive k of one output row
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (long long)gridDim.x * blockDim.x) {
      const int n = (int)(i % Y.N1), kc = (int)(i / Y.N1);
      const int nt = n >> 8, nin = n & 255;
      const int nrows = min(256, Y.N1 - nt * 256), rows_h = nrows >> 1;
      const int rank = nin >= rows_h ? 1 : 0, r = nin - rank * rows_h;
      float w[4], hi[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = kc * 4 + e;
        w[e] = (n < Y.out && k < Y.in) ? Ws[q.l_off[l] + (long long)n * q.l_ld[l] + k] : 0.0f;
        hi[e] = tc::tf32_hi(w[e]);
      }
      // N-tiles before this one hold 256 rows each: 2 * 256 * Kp floats per tile
      float* base = out + Y.w_off + (long long)nt * (2LL * 256 * Y.Kp) + (long long)rank * (2LL * rows_h * Y.Kp);
      *reinterpret_cast<float4*>(base + ((long long)kc * rows_h + r) * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      uint32_t* bf = reinterpret_cast<uint32_t*>(base + (long long)rows_h * Y.Kp) + ((long long)(kc >> 1) * rows_h + r) * 4 + (kc & 1) * 2;
      uint32_t* bl = bf + (long long)rows_h * Y.Kp / 2;
      bf[0] = tc::pack_bf16x2(w[0], w[1]); bf[1] = tc::pack_bf16x2(w[2], w[3]);
      bl[0] = tc::pack_bf16x2(w[0] - hi[0], w[1] - hi[1]); bl[1] = tc::pack_bf16x2(w[2] - hi[2], w[3] - hi[3]);
    }
  }
}

// ---- barriers ----------------------------------------------------------------------------------------------------------------
enum { TL_FULL = 0, TL_PFULL = TL_FULL + kTlStages, TL_EMPTY = TL_PFULL + kTlStages, TL_ACC_FULL = TL_EMPTY + kTlStages,
       TL_ACC_EMPTY = TL_ACC_FULL + 2, TL_ACT_READY = TL_ACC_EMPTY + 2, TL_COUNT };

struct TlSmemTail {
  uint64_t bars[TL_COUNT];
  uint32_t tmem_slot;
  float bias[2][256];
  float wlast[2][4][256];
  float ypart[128][4];
};

template <int kAct>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTlThreads, 1) fwd_tcl_kernel(const __grid_constant__ TlParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  TlSmemTail* T = reinterpret_cast<TlSmemTail*>(smem + kTlStages * kTlStageBytes);
  uint64_t* bars = T->bars;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = tl_ctarank();
  const bool leader = rank == 0;
  const int slot = blockIdx.x >> 1;
  const int g = blockIdx.y;
  const int s0 = g * p.Sg, s1 = min(p.S, s0 + p.Sg);
  const int pt0 = slot * p.tiles_per_slot, pt1 = min(p.n_tiles, pt0 + p.tiles_per_slot);
  const long long cta_lin = (long long)blockIdx.y * gridDim.x + blockIdx.x;
  float* scr = p.scratch + cta_lin * 2 * p.scratch_stride;

  if (tid == 0) {
    for (int i = 0; i < kTlStages; ++i) {
      mbar_init(&bars[TL_FULL + i], 1);
      mbar_init(&bars[TL_PFULL + i], 1);
      mbar_init(&bars[TL_EMPTY + i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[TL_ACC_FULL + i], 1);
      mbar_init(&bars[TL_ACC_EMPTY + i], 16);    // 8 epilogue warps of each CTA of the pair
    }
    mbar_init(&bars[TL_ACT_READY], 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&T->tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tl_cluster_sync();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = T->tmem_slot;
  const bool has_work = s1 > s0 && pt1 > pt0;

  if (has_work && warp == 0) {
    // ===================== loader: this CTA's 128 activation rows + its half of the weight rows ============================
    uint32_t pc = 0, act_cnt = 0;      // panels issued so far (ring position), hidden-layer hand-offs consumed so far
    for (int s = s0; s < s1; ++s) {
      const float* wi = p.wimg + (long long)s * p.wimg_stride;
      for (int pt = pt0; pt < pt1; ++pt) {
        for (int l = 0; l < p.Lh; ++l) {
          const TlLayer& Y = p.ly[l];
          const float* abase = l == 0 ? p.ximg + ((long long)pt * 2 + rank) * (2LL * 128 * Y.Kp) : scr + (long long)(l & 1) * p.scratch_stride;
          if (l > 0) {
            // the epilogue warps of THIS CTA have written this layer's A operand (all N-tiles of layer l - 1)
            mbar_wait(&bars[TL_ACT_READY], act_cnt & 1u);
            ++act_cnt;
          }
          for (int nt = 0; nt < Y.n_nt; ++nt) {
            const int nrows = min(256, Y.N1 - nt * 256), rows_h = nrows >> 1;
            const float* bbase = wi + Y.w_off + (long long)nt * (2LL * 256 * Y.Kp) + (long long)rank * (2LL * rows_h * Y.Kp);
            for (int kp = 0; kp < Y.n_kp; ++kp, ++pc) {
              const uint32_t st = pc % kTlStages, use = pc / kTlStages;
              mbar_wait(&bars[TL_EMPTY + st], (use & 1u) ^ 1u);
              if (tc::elect_one()) {
                unsigned char* sa = smem + st * kTlStageBytes;
                unsigned char* sb = sa + 32 * 1024;
                const uint32_t b32 = (uint32_t)(8 * rows_h * 16), b16 = (uint32_t)(4 * rows_h * 16);
                mbar_arrive_expect_tx(&bars[TL_FULL + st], 32u * 1024u + b32 + 2u * b16);
                bulk_g2s(sa, abase + (long long)kp * 8 * 128 * 4, 16u * 1024u, &bars[TL_FULL + st]);
                bulk_g2s(sa + 16 * 1024, abase + 128LL * Y.Kp + (long long)kp * 4 * 128 * 4, 8u * 1024u, &bars[TL_FULL + st]);
                bulk_g2s(sa + 24 * 1024, abase + 192LL * Y.Kp + (long long)kp * 4 * 128 * 4, 8u * 1024u, &bars[TL_FULL + st]);
                bulk_g2s(sb, bbase + (long long)kp * 8 * rows_h * 4, b32, &bars[TL_FULL + st]);
                bulk_g2s(sb + 16 * 1024, bbase + (long long)rows_h * Y.Kp + (long long)kp * 4 * rows_h * 4, b16, &bars[TL_FULL + st]);
                bulk_g2s(sb + 24 * 1024, bbase + (long long)rows_h * Y.Kp * 3 / 2 + (long long)kp * 4 * rows_h * 4, b16, &bars[TL_FULL + st]);
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (has_work && warp == 1) {
    if (!leader) {
      // ===================== peer CTA: forward "my part of the stage has landed" to the leader ==============================
      uint32_t pc = 0;
      for (int s = s0; s < s1; ++s)
        for (int pt = pt0; pt < pt1; ++pt)
          for (int l = 0; l < p.Lh; ++l)
            for (int nt = 0; nt < p.ly[l].n_nt; ++nt)
              for (int kp = 0; kp < p.ly[l].n_kp; ++kp, ++pc) {
                const uint32_t st = pc % kTlStages, use = pc / kTlStages;
                mbar_wait(&bars[TL_FULL + st], use & 1u);
                if (lane == 0) tl_arrive_remote(&bars[TL_PFULL + st], 0);
                __syncwarp();
              }
    } else {
      // ===================== MMA issuer (converged warp, one elected lane issues) ===========================================
      uint32_t pc = 0, tcnt = 0;
      const uint32_t s_base = smem_u32(smem);
      for (int s = s0; s < s1; ++s)
        for (int pt = pt0; pt < pt1; ++pt)
          for (int l = 0; l < p.Lh; ++l) {
            const TlLayer& Y = p.ly[l];
            for (int nt = 0; nt < Y.n_nt; ++nt, ++tcnt) {
              const int nrows = min(256, Y.N1 - nt * 256), rows_h = nrows >> 1;
              const uint32_t buf = tcnt & 1u;
              const uint32_t d = tmem_base + buf * 256u;
              const uint32_t id32 = tc::make_idesc_tf32(256, nrows), id16 = tc::make_idesc_bf16(256, nrows);
              mbar_wait(&bars[TL_ACC_EMPTY + buf], ((tcnt >> 1) & 1u) ^ 1u);   // the epilogues drained this accumulator (tile tcnt - 2)
              for (int kp = 0; kp < Y.n_kp; ++kp, ++pc) {
                const uint32_t st = pc % kTlStages, use = pc / kTlStages;
                mbar_wait(&bars[TL_FULL + st], use & 1u);
                mbar_wait(&bars[TL_PFULL + st], use & 1u);
                tc::fence_after_thread_sync();
                const uint32_t sa = s_base + st * kTlStageBytes, sb = sa + 32u * 1024u;
                const uint64_t a32 = tc::make_smem_desc(sa, 2048u, 128u), abf = tc::make_smem_desc(sa + 16u * 1024u, 2048u, 128u);
                const uint64_t abl = tc::make_smem_desc(sa + 24u * 1024u, 2048u, 128u);
                const uint64_t b32 = tc::make_smem_desc(sb, (uint32_t)rows_h * 16u, 128u);
                const uint64_t bbf = tc::make_smem_desc(sb + 16u * 1024u, (uint32_t)rows_h * 16u, 128u);
                const uint64_t bbl = tc::make_smem_desc(sb + 24u * 1024u, (uint32_t)rows_h * 16u, 128u);
                const uint64_t a_step = 2 * 128, b_step = (uint64_t)(2 * rows_h);   // 16-byte units per K = 8 (tf32) / K = 16 (bf16) step
                const uint32_t acc0 = kp > 0 ? 1u : 0u;
                const bool last = kp + 1 == Y.n_kp;
                if (tc::elect_one()) {
#pragma unroll
                  for (int k16 = 0; k16 < 2; ++k16) {
                    tl_mma_tf32(d, a32 + (uint64_t)(2 * k16) * a_step, b32 + (uint64_t)(2 * k16) * b_step, id32, k16 == 0 ? acc0 : 1u);
                    tl_mma_tf32(d, a32 + (uint64_t)(2 * k16 + 1) * a_step, b32 + (uint64_t)(2 * k16 + 1) * b_step, id32, 1u);
                    tl_mma_bf16(d, abl + (uint64_t)k16 * a_step, bbf + (uint64_t)k16 * b_step, id16, 1u);   // lo(a) . b
                    tl_mma_bf16(d, abf + (uint64_t)k16 * a_step, bbl + (uint64_t)k16 * b_step, id16, 1u);   // a . lo(b)
                  }
                  tl_commit(&bars[TL_EMPTY + st]);
                  if (last) tl_commit(&bars[TL_ACC_FULL + buf]);
                }
                __syncwarp();
              }
            }
          }
    }
  } else if (has_work && warp >= 2) {
    // ===================== epilogue warps: thread = one point (TMEM lane), half of the tile's columns =======================
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const int et = ew * 32 + lane;                 // 0..255 among the epilogue threads
    const long long M = p.N * p.O;
    double* st_mean = p.part ? p.part + (long long)g * 2 * M : nullptr;
    double* st_m2 = p.part ? st_mean + M : nullptr;
    uint32_t tcnt = 0;
    for (int s = s0; s < s1; ++s) {
      const float* Ws = p.W + (long long)s * p.w_stride;
      const double inv_cnt = 1.0 / (double)(s - s0 + 1);
      for (int pt = pt0; pt < pt1; ++pt) {
        const long long n = ((long long)pt * 2 + rank) * 128 + row;
        const bool in_range = n < p.N;
        float y[4] = {0.f, 0.f, 0.f, 0.f};
        for (int l = 0; l < p.Lh; ++l) {
          const TlLayer& Y = p.ly[l];
          const bool last_hidden = l + 1 == p.Lh;
          const int Kn = last_hidden ? 0 : p.ly[l + 1].Kp;                 // K extent of the images this layer writes
          float* ob = scr + (long long)((l + 1) & 1) * p.scratch_stride;
          for (int nt = 0; nt < Y.n_nt; ++nt, ++tcnt) {
            const int nrows = min(256, Y.N1 - nt * 256);
            const uint32_t buf = tcnt & 1u;
            // stage this tile's bias (and the last Linear's rows) in shared memory: one column per epilogue thread
            {
              const int c = nt * 256 + et;
              T->bias[buf][et] = c < Y.out ? __ldg(Ws + Y.b_off + (long long)c * Y.b_ld) : 0.0f;
              if (last_hidden)
                for (int o = 0; o < p.O; ++o) T->wlast[buf][o][et] = c < Y.out ? __ldg(Ws + p.last_off + (long long)o * p.last_ld + c) : 0.0f;
            }
            named_bar_sync(1, 256);
            const int nb16 = nrows >> 4;
            const int c_lo = half == 0 ? 0 : ((nb16 + 1) >> 1) * 16, c_hi = half == 0 ? ((nb16 + 1) >> 1) * 16 : nrows;
            mbar_wait(&bars[TL_ACC_FULL + buf], (tcnt >> 1) & 1u);
            tc::fence_after_thread_sync();
            const uint32_t d = tmem_base + buf * 256u + lane_sel;
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
              uint32_t v[16];
              tc::tmem_ld16_issue(d + (uint32_t)c0, v);
              tc::tmem_ld_wait(v);
              float h[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) h[j] = tl_act<kAct>(__uint_as_float(v[j]) + T->bias[buf][c0 + j]);
              if (!last_hidden) {
                // the three images of the next layer's A operand, K index = nt * 256 + c0 + j
                const int kg = nt * 256 + c0;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  float hi[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) hi[e] = tc::tf32_hi_fast(h[q4 * 4 + e]);
                  *reinterpret_cast<float4*>(ob + ((long long)((kg >> 2) + q4) * 128 + row) * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                }
                uint32_t* bf = reinterpret_cast<uint32_t*>(ob + 128LL * Kn) + ((long long)(kg >> 3) * 128 + row) * 4;
                uint32_t* bl = bf + 64LL * Kn;
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  uint4 ph, pl;
                  const float* hh = h + g8 * 8;
                  float lo[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) lo[e] = hh[e] - tc::tf32_hi_fast(hh[e]);
                  ph.x = tc::pack_bf16x2(hh[0], hh[1]); ph.y = tc::pack_bf16x2(hh[2], hh[3]);
                  ph.z = tc::pack_bf16x2(hh[4], hh[5]); ph.w = tc::pack_bf16x2(hh[6], hh[7]);
                  pl.x = tc::pack_bf16x2(lo[0], lo[1]); pl.y = tc::pack_bf16x2(lo[2], lo[3]);
                  pl.z = tc::pack_bf16x2(lo[4], lo[5]); pl.w = tc::pack_bf16x2(lo[6], lo[7]);
                  *reinterpret_cast<uint4*>(bf + (long long)g8 * 128 * 4) = ph;
                  *reinterpret_cast<uint4*>(bl + (long long)g8 * 128 * 4) = pl;
                }
              } else {
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                  if (o < p.O) {
                    float a = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) a = fmaf(T->wlast[buf][o][c0 + j], h[j], a);
                    y[o] += a;
                  }
                }
              }
            }
            // accumulator drained by this warp
            tc::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) tl_arrive_remote(&bars[TL_ACC_EMPTY + buf], 0);
          }
          if (!last_hidden) {
            // this warp's part of the next layer's A operand is in global memory: make it visible to the bulk-copy engine
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence_block();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[TL_ACT_READY]);
          }
        }
        // ---- last Linear's bias, the two column halves meet in shared memory, Welford update (as forward_tc.cu)
        if (half == 1) {
#pragma unroll
          for (int o = 0; o < 4; ++o) if (o < p.O) T->ypart[row][o] = y[o];
        }
        named_bar_sync(2 + quarter, 64);
        if (half == 0) {
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < p.O) y[o] += T->ypart[row][o] + __ldg(Ws + p.last_off + (long long)o * p.last_ld + p.H);
        }
        named_bar_sync(2 + quarter, 64);
        if (half == 0 && in_range) {
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            if (o < p.O) {
              const long long m = n * p.O + o;
              if (p.pred) p.pred[((long long)s * p.N + n) * p.O + o] = y[o];
              if (st_mean) {
                const double mu0 = s != s0 ? st_mean[m] : 0.0, q0 = s != s0 ? st_m2[m] : 0.0;
                const double dlt = (double)y[o] - mu0;
                const double mu = mu0 + dlt * inv_cnt;
                st_mean[m] = mu;
                st_m2[m] = q0 + dlt * ((double)y[o] - mu);
              }
            }
          }
        }
      }
    }
  }

  tc::fence_before_thread_sync();
  __syncthreads();
  tl_cluster_sync();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
static inline long long rup(long long v, long long m) { return (v + m - 1) / m * m; }

struct TlPlan {
  TlParams p;
  TlPrepW q;
  long long ximg_bytes, wimg_bytes, scratch_bytes, part_bytes;
  int slots;
  size_t smem;
  long long n_tiles128;
};

int finalize_moments(const double* part, const int*, int G, int Sg, long long S, long long M, float* mean, float* var,
                     double* part_mean, double* part_m2, cudaStream_t st);

// 0 = ok, 1 = shape not supported (message set), < 0 error
static int plan_tcl(const BnnForwardArgs* a, TlPlan* plan) {
  Layout L;
  if (make_layout(&a->desc, &L)) return -1;
  if (a->precision != BNN_PREC_TF32_BF16) { set_error("deep / wide tensor-core forward: only BNN_PREC_TF32_BF16"); return 1; }
  if (L.n_lin < 2) { set_error("deep / wide tensor-core forward needs at least one hidden layer"); return 1; }
  if (a->mask.mode != BNN_MASK_NONE) { set_error("deep / wide tensor-core forward: dropout masks are not supported (use BNN_PREC_FP32)"); return 1; }
  if (a->w_stride == 0) { set_error("deep / wide tensor-core forward: per-sample banks only"); return 1; }
  const int O = L.out[L.n_lin - 1];
  if (O > 4) { set_error("deep / wide tensor-core forward: nout <= 4"); return 1; }
  TlParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  memset(&plan->q, 0, sizeof(plan->q));
  p.Lh = L.n_lin - 1; p.act = L.act; p.O = O; p.D = L.in[0];
  long long woff = 0, kp_max = 32;
  for (int l = 0; l < p.Lh; ++l) {
    TlLayer& Y = p.ly[l];
    Y.in = L.in[l]; Y.out = L.out[l];
    Y.Kp = (int)rup(Y.in, kTlKP); Y.N1 = (int)rup(Y.out, 16);
    Y.n_nt = (Y.N1 + 255) / 256; Y.n_kp = Y.Kp / kTlKP;
    if (Y.n_nt > kTlMaxNT) { set_error("deep / wide tensor-core forward: hidden width <= 1024"); return 1; }
    Y.w_off = woff;
    woff += (long long)Y.n_nt * 2 * 256 * Y.Kp;
    Y.b_off = L.off[l] + L.in[l]; Y.b_ld = L.ld[l];
    if (Y.Kp > kp_max) kp_max = Y.Kp;
    plan->q.l_off[l] = L.off[l]; plan->q.l_ld[l] = L.ld[l];
  }
  for (int l = 0; l + 1 < p.Lh; ++l)
    if (rup(p.ly[l].out, kTlKP) != p.ly[l + 1].Kp) { set_error("internal: layer widths"); return -1; }
  p.H = L.in[L.n_lin - 1];
  p.last_off = L.off[L.n_lin - 1]; p.last_ld = L.ld[L.n_lin - 1];
  p.wimg_stride = woff;
  p.W = a->W; p.w_stride = a->w_stride;
  p.N = a->N; p.S = (int)a->S; p.pred = a->pred;
  p.n_tiles = (int)((a->N + 255) / 256);
  plan->n_tiles128 = (long long)p.n_tiles * 2;
  const int max_slots = sm_count() / 2;
  int slots = p.n_tiles < max_slots ? p.n_tiles : max_slots;
  p.tiles_per_slot = (p.n_tiles + slots - 1) / slots;
  slots = (p.n_tiles + p.tiles_per_slot - 1) / p.tiles_per_slot;
  plan->slots = slots;
  long long G = 1;
  if (slots < max_slots) G = max_slots / slots;
  if (G > a->S) G = a->S;
  if (G < 1) G = 1;
  if (G > 65535) G = 65535;
  const long long Sg = (a->S + G - 1) / G;
  G = (a->S + Sg - 1) / Sg;
  p.G = (int)G; p.Sg = (int)Sg;
  p.scratch_stride = 2LL * 128 * kp_max;
  plan->ximg_bytes = rup(plan->n_tiles128 * 2 * 128 * p.ly[0].Kp * 4, 256);
  plan->wimg_bytes = rup((long long)a->S * p.wimg_stride * 4, 256);
  plan->scratch_bytes = rup((long long)slots * 2 * G * 2 * p.scratch_stride * 4, 256);
  plan->part_bytes = (long long)G * 2 * a->N * O * (long long)sizeof(double);
  plan->smem = (size_t)kTlStages * kTlStageBytes + sizeof(TlSmemTail);
  plan->q.Lh = p.Lh;
  for (int l = 0; l < p.Lh; ++l) plan->q.ly[l] = p.ly[l];
  plan->q.w_stride = a->w_stride; plan->q.wimg_stride = p.wimg_stride;
  return 0;
}

long long forward_tcl_workspace_bytes(const BnnForwardArgs* a) {
  TlPlan plan;
  if (plan_tcl(a, &plan)) return -1;
  return plan.ximg_bytes + plan.wimg_bytes + plan.scratch_bytes + plan.part_bytes;
}

int forward_tcl_supported(const BnnForwardArgs* a) {
  TlPlan plan;
  return plan_tcl(a, &plan) == 0;
}

int forward_tcl(const BnnForwardArgs* a, cudaStream_t st) {
  TlPlan plan;
  if (plan_tcl(a, &plan)) return -1;
  TlParams& p = plan.p;
  const long long need = plan.ximg_bytes + plan.wimg_bytes + plan.scratch_bytes + plan.part_bytes;
  if (!a->workspace || a->workspace_bytes < need)
    BNN_FAIL("workspace too small: need %lld bytes, got %lld", need, (long long)a->workspace_bytes);
  char* ws = reinterpret_cast<char*>(a->workspace);
  float* ximg = reinterpret_cast<float*>(ws);
  float* wimg = reinterpret_cast<float*>(ws + plan.ximg_bytes);
  float* scratch = reinterpret_cast<float*>(ws + plan.ximg_bytes + plan.wimg_bytes);
  const bool want_moments = a->mean || a->var || a->part_mean || a->part_m2;
  p.part = want_moments ? reinterpret_cast<double*>(ws + plan.ximg_bytes + plan.wimg_bytes + plan.scratch_bytes) : nullptr;
  p.ximg = ximg; p.wimg = wimg; p.scratch = scratch;
  // pad columns of the activation images (K beyond a layer's width) are read as operands but never written: zero them once
  BNN_CUDA(cudaMemsetAsync(scratch, 0, (size_t)plan.scratch_bytes, st));
  {
    const long long quads = plan.n_tiles128 * 128 * (p.ly[0].Kp / 4);
    long long blocks = (quads + 255) / 256;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    tl_prep_x_kernel<<<(unsigned)blocks, 256, 0, st>>>(a->X, a->N, p.D, p.ly[0].Kp, plan.n_tiles128, ximg);
    BNN_CUDA(cudaGetLastError());
    for (long long s0 = 0; s0 < a->S; s0 += 65535) {
      const long long ns = a->S - s0 < 65535 ? a->S - s0 : 65535;
      dim3 grid(64, (unsigned)ns);
      tl_prep_w_kernel<<<grid, 256, 0, st>>>(plan.q, a->W + s0 * a->w_stride, wimg + s0 * p.wimg_stride);
      BNN_CUDA(cudaGetLastError());
    }
  }
  dim3 grid((unsigned)(plan.slots * 2), (unsigned)p.G);
  auto launch = [&](auto kern) -> int {
    BNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    kern<<<grid, kTlThreads, plan.smem, st>>>(p);
    BNN_CUDA(cudaGetLastError());
    return 0;
  };
  int rc = 0;
  switch (p.act) {
    case BNN_ACT_RELU: rc = launch(fwd_tcl_kernel<BNN_ACT_RELU>); break;
    case BNN_ACT_TANH: rc = launch(fwd_tcl_kernel<BNN_ACT_TANH>); break;
    case BNN_ACT_SIGMOID: rc = launch(fwd_tcl_kernel<BNN_ACT_SIGMOID>); break;
    default: rc = launch(fwd_tcl_kernel<BNN_ACT_IDENTITY>); break;
  }
  if (rc) return rc;
  if (want_moments)
    return finalize_moments(p.part, nullptr, p.G, p.Sg, a->S, a->N * p.O, a->mean, a->var, a->part_mean, a->part_m2, st);
  return 0;
}

}  // namespace bnn
