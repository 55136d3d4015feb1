// Kernel (a)+(e), tensor-core variant: fused S x N batched MLP forward on tcgen05 (5th-gen tensor cores),
// accumulators in TMEM, FP32-grade accuracy through error-compensated products.  With a = a_h + a_l, a_h = tf32(a):
//     a*b ~= a_h*b_h + a_l*b + a*b_l
//   BNN_PREC_TF32X3   : all three terms as TF32 MMAs (kind::tf32, K = 8)
//   BNN_PREC_TF32_BF16: main term TF32, the two (2^-11 small) correction terms as BF16 MMAs (kind::f16, K = 16) on
//                       bf16(a_l)*bf16(b) and bf16(a)*bf16(b_l) -- 4 instead of 6 bf16-equivalent passes per MAC
//   BNN_PREC_TF32     : a single TF32 pass (looser bound, opt-in).
// Replaces the same reference code as forward_simt.cu (the `for i in range(S): pred[i] = nns[i](X)` loops +
// BNN.py:36-37) for the wide configs where the layers are real contractions (M = 128/256 points, N = hidden width,
// K = previous width).
//
// Shape family: dims = [D, H1, H2, O] with D <= 64, H1, H2 <= 256, O <= 4 (the last Linear is a dot product in
// the epilogue).  One CTA (kPair = 1) or one CTA pair on two SMs (kPair = 2, cta_group::2, M = 256) owns a range
// of 128-point tiles and walks the posterior samples in the outer loop:
//   per sample : the sample's weights, pre-split into UMMA operand images by tc_prep_kernel, are bulk-copied (TMA bulk
//                engine) into shared memory ONCE and stay resident for all tiles (pair mode: each CTA holds half of
//                the N rows of every B operand, the tensor core reads both halves);
//   per tile   : X tile -> A1 images (smem) -> MMA1 -> D1[TMEM] -> epilogue-1 (bias, activation, split) writes the
//                layer-2 A operand stage by stage BACK INTO TENSOR MEMORY (tcgen05.st; a ring of 32-column stages
//                behind D1 and D2) -> MMA2 reads A from TMEM (TS form) and B2 from shared memory into D2[TMEM] ->
//                epilogue-2 (bias, activation, dot with w3) -> y -> FP64 Welford update of the per-point state
//                (global, owned by this CTA; finalised by finalize_moments()).
// Roles (544 threads): warps 0-7 epilogue-1 (+ A1), warps 8-15 epilogue-2 + reduction, warp 16 weight loader + MMA issuer.
#include <stdlib.h>

#include "common.cuh"
#include "philox.cuh"
#include "tc_common.cuh"
#include <type_traits>

namespace bnn {

// Warp roles.  BNN_TC_SETS = number of epilogue-1 warp sets (4 warps each, one per TMEM lane quarter; a set takes every
// BNN_TC_SETS-th 16-column stage of a tile).  2: 8 + 8 + 1 = 17 warps (allocated like 20: 96 registers per thread).
// 3: 12 + 8 + 1 (+ 3 idle warps completing the last warpgroup) = 24 warps launched at 80 registers per thread and rebalanced
// with setmaxnreg: the MMA / loader warpgroup gives registers back, the epilogue-2 warpgroups (which keep a 32-column TMEM
// batch, the running sums and the FP64 Welford state in registers) take them.
#ifndef BNN_TC_SETS
#define BNN_TC_SETS 2
#endif
constexpr int kTcSets = BNN_TC_SETS;
static_assert(kTcSets == 2 || kTcSets == 3, "BNN_TC_SETS must be 2 or 3");
constexpr int kTcE1Warps = 4 * kTcSets;
constexpr int kTcE1Threads = 32 * kTcE1Warps;
constexpr int kTcMmaWarp = kTcE1Warps + 8;
constexpr int kTcThreads = kTcSets == 2 ? 544 : 768;
#ifndef BNN_TC_REG_E2
#define BNN_TC_REG_E2 88
#endif
#ifndef BNN_TC_REG_MMA
#define BNN_TC_REG_MMA 40
#endif
#ifndef BNN_TC_REG_E1
#define BNN_TC_REG_E1 88
#endif
#ifndef BNN_TC_REG_SLACK
#define BNN_TC_REG_SLACK 0      /* registers of the SM the launch left unallocated (65536 - 768 x 80 = 4096), if the pool has them */
#endif
// registers released by the MMA warpgroup must cover what the epilogue warpgroups request (setmaxnreg.inc blocks otherwise)
static_assert(kTcSets == 2 || 128 * (80 - BNN_TC_REG_MMA) >= 384 * (BNN_TC_REG_E1 - 80) + 256 * (BNN_TC_REG_E2 - 80) - BNN_TC_REG_SLACK, "register budget");
template <int kRegs> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
constexpr int kTcRing = 6;            // max A2 ring stages
constexpr int kTcStageCols = 32;       // one A2 stage = 16 K-columns (two K = 8 MMA steps) in tensor memory: hi[16] then lo[16]

struct TcParams {
  int D, H1, H2, O, act;
  int Dp, N1, N2, K2;      // padded: K of layer 1 (mult of 8; 16 in mixed mode), N of layers 1/2 (mult of 16), K of layer 2 (<= N1)
  int ring;                // A2 ring depth (2..kTcRing) in 16-column stages (tensor memory, after D1 and D2)
  int passes;              // 3 (tf32x3), 2 (tf32 + two bf16 correction terms) or 1 (tf32)
  long long part_stride;   // floats per (sample, part) of the operand image
  long long img_stride;    // floats per sample of the operand image
  int vec_len;             // floats: b1[N1] b2[N2] w3[O*N2] b3[4]
  const float* img;        // [S][img_stride]
  const float* X;
  long long N;
  int S, G, Sg;
  int tiles_per_slot, n_tiles;   // pair-tiles (128 * kPair points each)
  float* pred;
  double* part;            // [G][2][M]  Welford state = shard moments
  unsigned int* progress;  // [G][slots] samples started per CTA pair (pacing, see the loader warp), zero before launch; may be null
  int pace;                // samples a pair may run ahead of the slowest pair of its group
  int tmem_cols;
  // per-sample input-unit masks (dropout family): m_off[l] = offset of Linear l's input mask in the mask vector, -1 = none
  int mask_mode, mask_len, m_off[3];
  const float* mask;
  long long mask_stride;
  float p_drop[3], temperature;
  unsigned long long seed;
  long long sample_offset;
  int shared_w;            // 1: one network for all samples (w_stride == 0): operands are loaded once
  unsigned long long* trace;   // debug timeline (BNN_TC_TRACE), null in production
  int trace_mask;              // BNN_TC_TRACE value: bit 0 MMA per-stage events, bit 1 WG-A, bit 2 WG-B (tile ends always)
  // shared memory offsets (bytes)
  int sm_b1, sm_b2, sm_vec, sm_a1, sm_bar, sm_ypart, sm_mask;
};

// ---- operand image builder -------------------------------------------------------------------------
struct TcPrepParams {
  int D, H1, H2, O, Dp, N1, N2, K2, kpair;
  int ld0, ld1, ld2;
  long long off0, off1, off2, w_stride, part_stride, img_stride;
  int vec_len;
  int mixed;   // second half of each operand region = two BF16 images (pairs per word) instead of the TF32 lo image
  long long S; // networks to convert (1 for a shared network)
};

__global__ void __launch_bounds__(256) tc_prep_kernel(const __grid_constant__ TcPrepParams q, const float* __restrict__ W,
                                                      float* __restrict__ img) {
  for (long long s = blockIdx.y; s < q.S; s += gridDim.y) {   // grid.y is clamped to 65535: stride over the samples
  const float* Ws = W + s * q.w_stride;
  float* out = img + s * q.img_stride;
  // all per-sample indices fit 32 bits (one image is < 2^20 floats): 32-bit divisions, the 64-bit ones dominated this kernel
  const int N1h = q.N1 / q.kpair, N2h = q.N2 / q.kpair;
  const int b1_sz = q.Dp * N1h, b2_sz = q.K2 * N2h;
  const int total = (int)q.img_stride, part_stride = (int)q.part_stride;
  const int off0 = (int)q.off0, off1 = (int)q.off1, off2 = (int)q.off2;
  for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < total; i += (int)(gridDim.x * blockDim.x)) {
    float v = 0.0f;
    int r = i;
    if (r < part_stride * q.kpair) {
      const int part = r >= part_stride ? 1 : 0;      // kpair <= 2
      r -= part * part_stride;
      // regions: [B1 hi][B1 lo][B2 hi][B2 lo], each [kc][n][4]
      int layer, hl;
      if (r < 2 * b1_sz) { layer = 0; hl = r >= b1_sz; r -= hl * b1_sz; }
      else { r -= 2 * b1_sz; layer = 1; hl = r >= b2_sz; r -= hl * b2_sz; }
      const int rows = layer == 0 ? N1h : N2h;
      auto weight = [&](int ng, int k) {
        float w = 0.0f;
        if (layer == 0) { if (ng < q.H1 && k < q.D) w = Ws[off0 + ng * q.ld0 + k]; }
        else            { if (ng < q.H2 && k < q.H1) w = Ws[off1 + ng * q.ld1 + k]; }
        return w;
      };
      if (hl && q.mixed) {
        // BF16 images [kc8][n][8 x bf16]: bf16(w) in the first half of the region, bf16(w - tf32(w)) in the second
        const int half_sz = (layer == 0 ? b1_sz : b2_sz) >> 1;
        const int lo_img = r >= half_sz;
        r -= lo_img * half_sz;
        const int e = r & 3;
        const int g = r >> 2;
        const int kc8 = g / rows, n = g - kc8 * rows;
        const int k = kc8 * 8 + 2 * e, ng = part * rows + n;
        const float w0 = weight(ng, k), w1 = weight(ng, k + 1);
        const uint32_t pk = lo_img ? tc::pack_bf16x2(w0 - tc::tf32_hi(w0), w1 - tc::tf32_hi(w1)) : tc::pack_bf16x2(w0, w1);
        v = __uint_as_float(pk);
      } else {
        const int e = r & 3;
        const int g = r >> 2;
        const int kc = g / rows, n = g - kc * rows;
        const float w = weight(part * rows + n, kc * 4 + e);
        const float hi = tc::tf32_hi(w);
        v = hl ? (w - hi) : hi;
      }
    } else {
      int j = r - part_stride * q.kpair;
      if (j < q.N1) { if (j < q.H1) v = Ws[off0 + j * q.ld0 + q.D]; }                      // b1
      else if ((j -= q.N1) < q.N2) { if (j < q.H2) v = Ws[off1 + j * q.ld1 + q.H1]; }      // b2
      else if ((j -= q.N2) < q.O * q.N2) {                                                 // w3[o][j]
        const int o = j / q.N2, c = j - o * q.N2;
        if (c < q.H2) v = Ws[off2 + o * q.ld2 + c];
      } else if ((j -= q.O * q.N2) < q.O) v = Ws[off2 + j * q.ld2 + q.H2];                 // b3
    }
    out[i] = v;
  }
  }
}

// ---- cluster helpers (pair mode) ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"   // default .release.cta: a cluster-scope release costs a MEMBAR.ALL.GPU per hand-off
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

template <int kPair>
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kPair == 1) {
    tc::mma_tf32_ss(tmem_d, adesc, bdesc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// same with the A operand in tensor memory (each CTA of a pair supplies its own 128 rows at the same TMEM address)
template <int kPair>
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kPair == 1) {
    tc::mma_tf32_ts(tmem_d, tmem_a, bdesc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// kind::f16 (BF16 operands, K = 16) variants for the correction terms of the mixed scheme
template <int kPair>
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kPair == 1) {
    tc::mma_bf16_ss(tmem_d, adesc, bdesc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
template <int kPair>
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kPair == 1) {
    tc::mma_bf16_ts(tmem_d, tmem_a, bdesc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// MMA completion -> arrival on the barrier at this offset in every CTA of the pair
template <int kPair>
__device__ __forceinline__ void mma_commit_all(uint64_t* bar) {
  if constexpr (kPair == 1) {
    tc::mma_commit(bar);
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
  }
}
template <int kPair>
__device__ __forceinline__ void tmem_alloc_n(uint32_t* dst, uint32_t ncols) {
  if constexpr (kPair == 1) {
    tc::tmem_alloc(dst, ncols);
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int kPair>
__device__ __forceinline__ void tmem_dealloc_n(uint32_t addr, uint32_t ncols) {
  if constexpr (kPair == 1) tc::tmem_dealloc(addr, ncols);
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// debug timeline: CTA (0,0) only; each traced warp owns 4096 (tag, clock) slots
#ifdef BNN_TC_TRACE_BUILD
__device__ __forceinline__ void trace_ev(unsigned long long* tr, int who, uint32_t& n, uint32_t tag, uint32_t it) {
  if (tr && n < 4096) {
    tr[((size_t)who * 4096 + n) * 2] = ((unsigned long long)tag << 32) | it;
    tr[((size_t)who * 4096 + n) * 2 + 1] = (unsigned long long)clock64();
    ++n;
  }
}
#else
// production build: the timeline probes compile to nothing (each one costs the single-threaded MMA issuer ~70 cycles)
__device__ __forceinline__ void trace_ev(unsigned long long*, int, uint32_t&, uint32_t, uint32_t) {}
#endif

// keep-mask value of input unit j_global (index into the concatenated mask vector) of `layer` for global sample s:
// identical streams to forward_simt.cu / bnn_dropout_masks, so all three agree bit for bit
__device__ __forceinline__ float tc_mask_value(const TcParams& p, long long s_global, int s_local, int layer, int j_global) {
  if (p.mask_mode == BNN_MASK_EXTERNAL) return __ldg(p.mask + (long long)s_local * p.mask_stride + j_global);
  const uint32_t tag = p.mask_mode == BNN_MASK_BERNOULLI ? STREAM_BERNOULLI : STREAM_CONCRETE;
  const u32x4 blk = philox_block(p.seed, (uint32_t)s_global, tag, (uint64_t)(j_global >> 2));
  const float u = u24(pick_word(blk, j_global & 3));
  return p.mask_mode == BNN_MASK_BERNOULLI ? bernoulli_keep(u, p.p_drop[layer]) : concrete_keep(u, p.p_drop[layer], p.temperature);
}

// barrier slots in shared memory
enum { BAR_W_FULL = 0, BAR_A1_FULL, BAR_D1_FULL, BAR_D2_FULL, BAR_D2_EMPTY, BAR_SAMPLE_DONE, BAR_VEC_FREE, BAR_A2_FULL,
       BAR_A2_EMPTY = BAR_A2_FULL + kTcRing,
       BAR_COUNT = BAR_A2_EMPTY + kTcRing };

// one warp signals "my part is done" to the MMA issuer (leader CTA): writers fence their generic-proxy
// stores for the async proxy, the warp converges, one lane arrives
template <int kPair>
__device__ __forceinline__ void warp_signal_leader(uint64_t* bar, int lane) {
  fence_proxy_async();
  tc::fence_before_thread_sync();
  __syncwarp();
  if (lane == 0) {
    if constexpr (kPair == 1) mbar_arrive(bar);
    else mbar_arrive_cluster(bar, 0);
  }
}

// same for data handed over in tensor memory (tcgen05.st completed by the caller): no async-proxy fence needed
template <int kPair>
__device__ __forceinline__ void warp_signal_tmem(uint64_t* bar, int lane) {
  tc::fence_before_thread_sync();
  __syncwarp();
  if (lane == 0) {
    if constexpr (kPair == 1) mbar_arrive(bar);
    else mbar_arrive_cluster(bar, 0);
  }
}

// Activations of the tensor-core path.  tanh / sigmoid go through MUFU.EX2 + MUFU.RCP (5 / 4 instructions; libdevice's
// tanhf is ~18 with both of its branches predicated, and the epilogues are instruction-bound): absolute error <= 2e-7 on
// values in [-1, 1], the same order as the 3xTF32 / TF32+BF16 products feeding them (parity tests: 1e-5 max-normalised).
// The FP32 SIMT path keeps tanhf / expf.
template <int kAct>
__device__ __forceinline__ float act_ct(float v) {
  if constexpr (kAct == BNN_ACT_RELU) return fmaxf(v, 0.0f);
  else if constexpr (kAct == BNN_ACT_TANH) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 2.8853900817779268f));   // exp(2 v)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);                                                   // 1 - 2 / (exp(2 v) + 1)
  } else if constexpr (kAct == BNN_ACT_SIGMOID) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));  // exp(-v)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return r;
  } else return v;
}

// Roles: warps 0-7 (WG-A) run epilogue-1, two warps per TMEM lane quarter taking alternate 16-column batches
// (warps 0-3 also build A1); warps 8-15 (WG-B) run epilogue-2 (two per lane quarter, each reducing half of the D2
// columns; partial outputs meet in shared memory) + the Welford update, overlapping layer 1 / epilogue-1 of tile t+1;
// warp 16 = weight loader + MMA issuer.
// A warp may only touch TMEM lanes [32 * (warp % 4), +32).
//
// Hand-off protocol (all mbarriers; "leader" = CTA 0 of the pair, the only one that issues MMAs):
//   W_FULL       tx barrier: this sample's operand images + bias vector have landed in THIS CTA
//   A1_FULL      4 x kPair warp arrivals on the leader: A1 of the tile is written (and W_FULL was seen in that CTA)
//   D1_FULL      commit (multicast): layer-1 MMAs done -> epilogue-1 may read D1, A1 may be rebuilt
//   A2_FULL[st]  4 x kPair warp arrivals on the leader: stage st (16 K-columns of the layer-2 A operand) is written
//   A2_EMPTY[st] commit (multicast): the MMAs reading stage st are done
//   D2_FULL      commit (multicast): layer-2 MMAs of the tile done -> epilogue-2 may read D2
//   D2_EMPTY     8 x kPair warp arrivals on the leader: epilogue-2 drained D2
//   SAMPLE_DONE  commit (multicast) after the last tile of a sample: B1/B2 may be overwritten
//   VEC_FREE     8 local warp arrivals: epilogue-2 is done with this sample's bias vector (single-buffered)
template <int kPair, int kAct>
__global__ void __launch_bounds__(kTcThreads, 1) fwd_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.sm_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  const uint32_t s_b1 = smem_u32(smem + p.sm_b1), s_b2 = smem_u32(smem + p.sm_b2);
  const uint32_t s_a1 = smem_u32(smem + p.sm_a1);
  const uint32_t s_vec = smem_u32(smem + p.sm_vec);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = kPair == 1 ? 0u : cluster_ctarank();
  const bool leader = rank == 0;
  const int slot = blockIdx.x / kPair;
  const int g = blockIdx.y;
  const int s0 = g * p.Sg, s1 = min(p.S, s0 + p.Sg);
  const int pt0 = slot * p.tiles_per_slot, pt1 = min(p.n_tiles, pt0 + p.tiles_per_slot);
  const int N1h = p.N1 / kPair, N2h = p.N2 / kPair;
  const int n_k1 = p.Dp >> 3, n_k2 = p.K2 >> 3;
  const int n_b2 = (n_k2 + 1) >> 1;   // 16-column batches (= ring stages) per tile; the last may hold one K-step
  const uint32_t b1_hl_bytes = (uint32_t)p.Dp * N1h * 4, b2_hl_bytes = (uint32_t)p.K2 * N2h * 4;
  const uint32_t a1_hl_bytes = (uint32_t)p.Dp * 128 * 4;
  const int vec_pad = (p.vec_len + 3) & ~3;

  if (tid == 0) {
    mbar_init(&bars[BAR_W_FULL], 1);
    mbar_init(&bars[BAR_A1_FULL], 4 * kPair);
    mbar_init(&bars[BAR_D1_FULL], 1);
    mbar_init(&bars[BAR_D2_FULL], 1);
    mbar_init(&bars[BAR_D2_EMPTY], 8 * kPair);
    mbar_init(&bars[BAR_SAMPLE_DONE], 1);
    mbar_init(&bars[BAR_VEC_FREE], 8);
    for (int i = 0; i < kTcRing; ++i) {
      mbar_init(&bars[BAR_A2_FULL + i], 4 * kPair);
      mbar_init(&bars[BAR_A2_EMPTY + i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kTcMmaWarp) tmem_alloc_n<kPair>(tmem_slot, (uint32_t)p.tmem_cols);
  tc::fence_before_thread_sync();
  __syncthreads();
  if constexpr (kPair == 2) cluster_sync_all();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + (uint32_t)p.N1, tmem_a2 = tmem_d2 + (uint32_t)p.N2;

  if (warp >= kTcMmaWarp) {
    if constexpr (kTcSets == 3) setmaxnreg_dec<BNN_TC_REG_MMA>();   // whole warpgroup (warps 20-23); 21-23 have no other work
  }
  if (warp == kTcMmaWarp) {
    // ===================== weight loader + MMA issuer =========================================================
    // The whole warp runs this loop CONVERGED with warp-uniform variables; only the tcgen05 / bulk-copy issue is
    // done by one elected lane.  (Running the loop inside `if (lane == 0)` makes ptxas treat it as divergent code
    // and wrap every uniform-datapath instruction -- UTCHMMA, UTCBAR -- in an ELECT/BRA.U.ANY loop with R2UR
    // transfers: ~90 instructions and ~700 cycles per K-step, which starved the tensor pipe.)
    const uint32_t idesc1 = tc::make_idesc_tf32(128 * kPair, p.N1), idesc2 = tc::make_idesc_tf32(128 * kPair, p.N2);
    uint32_t tile_cnt = 0, tn = 0, ring_st = 0, ring_par = 0;
    unsigned long long* tr = (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) ? p.trace : nullptr;
    unsigned long long* trs = (p.trace_mask & 1) ? tr : nullptr;   // per-stage events (each costs this warp ~70 cycles)
    const uint64_t a1_desc0 = tc::make_smem_desc(s_a1, 2048u, 128u);
    const uint64_t b1_desc0 = tc::make_smem_desc(s_b1, (uint32_t)N1h * 16u, 128u);
    const uint64_t b2_desc0 = tc::make_smem_desc(s_b2, (uint32_t)N2h * 16u, 128u);
    const uint64_t b1_step = (uint64_t)(2 * N1h), b2_step = (uint64_t)(2 * N2h);   // 16-byte units per K = 8 step
    const uint64_t a_step = (uint64_t)(2 * 128);
    const bool three = p.passes == 3, mixed = p.passes == 2;
    const uint32_t idesc1h = tc::make_idesc_bf16(128 * kPair, p.N1), idesc2h = tc::make_idesc_bf16(128 * kPair, p.N2);
    trace_ev(tr, 0, tn, 20, 0);
    bool pace_on = true;
    for (int s = s0; s < s1; ++s) {
      const uint32_t spar = (uint32_t)(s - s0) & 1u;
      const bool reload = !p.shared_w || s == s0;   // a shared network is loaded once and stays resident
      // all MMAs of the previous sample are complete in both CTAs -> B1 / B2 may be overwritten
      if (reload && s > s0) mbar_wait(&bars[BAR_SAMPLE_DONE], spar ^ 1u);
      const long long s_img = p.shared_w ? 0 : s;
      const float* src = p.img + s_img * p.img_stride + (long long)rank * p.part_stride;
      const uint32_t b1_bytes = (three || mixed) ? 2 * b1_hl_bytes : b1_hl_bytes;
      const uint32_t b2_bytes = (three || mixed) ? 2 * b2_hl_bytes : b2_hl_bytes;
      // Pacing.  Pairs own 52 or 53 tiles at cfg 4, so over thousands of samples the fast pairs drift hundreds of samples ahead
      // and every pair ends up fetching "its" copy of each sample's operand image from DRAM (ncu: 44 MB of DRAM traffic per
      // sample at S = 10,000 vs 5 MB at S = 2,048).  A pair therefore does not start sample s before the slowest pair of its
      // group has started sample s - pace: the group's image loads stay within what L2 holds.  Progress counters are plain
      // relaxed stores / loads; a wait gives up for good after ~2 ms (not all CTAs resident: no deadlock, just no pacing).
      if (reload && p.progress != nullptr) {
        unsigned int* prog = p.progress + (long long)g * gridDim.x / kPair;
        const unsigned int rel = (unsigned int)(s - s0);
        if (leader && lane == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(prog + slot), "r"(rel + 1u) : "memory");
        if (pace_on && rel > (unsigned int)p.pace) {
          const unsigned int need = rel - (unsigned int)p.pace;
          const int npairs = (int)(gridDim.x / kPair);
          const long long t_start = clock64();
          for (;;) {
            unsigned int mn = 0xffffffffu;
            for (int j = lane; j < npairs; j += 32) {
              unsigned int v;
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(prog + j) : "memory");
              mn = min(mn, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            if (mn >= need) break;
            const int late = __shfl_sync(0xffffffffu, (int)(clock64() - t_start > 4000000LL), 0);   // warp-uniform decision
            if (late) { pace_on = false; break; }
          }
        }
      }
      if (reload && tc::elect_one()) {
        mbar_arrive_expect_tx(&bars[BAR_W_FULL], b1_bytes + b2_bytes + (uint32_t)vec_pad * 4);
        bulk_g2s(smem + p.sm_b1, src, b1_bytes, &bars[BAR_W_FULL]);
        bulk_g2s(smem + p.sm_b2, src + 2 * (long long)(b1_hl_bytes / 4), b2_bytes, &bars[BAR_W_FULL]);
      }
      __syncwarp();
      // the bias vector is single-buffered: epilogue-2 of the previous sample must be done with it
      if (reload && s > s0) mbar_wait(&bars[BAR_VEC_FREE], spar ^ 1u);
      if (reload && tc::elect_one())
        bulk_g2s(smem + p.sm_vec, p.img + s_img * p.img_stride + p.part_stride * kPair, (uint32_t)vec_pad * 4,
                 &bars[BAR_W_FULL]);
      __syncwarp();
      if (!leader) continue;   // the peer CTA only loads; its tensor-core work is issued by the leader
      if (reload) mbar_wait(&bars[BAR_W_FULL], spar);
      for (int pt = pt0; pt < pt1; ++pt, ++tile_cnt) {
        const uint32_t tpar = tile_cnt & 1u;
        // ---- layer 1: D1 = A1 * B1^T
        mbar_wait(&bars[BAR_A1_FULL], tpar);
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
          uint64_t a_hi = a1_desc0, b_hi = b1_desc0;
          const uint64_t a_lo_off = (uint64_t)(a1_hl_bytes >> 4), b_lo_off = (uint64_t)(b1_hl_bytes >> 4);
          for (int ks = 0; ks < n_k1; ++ks) {
            mma_tf32<kPair>(tmem_d1, a_hi, b_hi, idesc1, ks > 0);
            if (three) {
              mma_tf32<kPair>(tmem_d1, a_hi + a_lo_off, b_hi, idesc1, 1);
              mma_tf32<kPair>(tmem_d1, a_hi, b_hi + b_lo_off, idesc1, 1);
            }
            if (mixed && (ks & 1)) {
              // correction terms of this K = 16 block from the BF16 images (bf16(a) | bf16(a_lo) behind the TF32 image)
              const uint64_t a16 = a1_desc0 + a_lo_off + (uint64_t)(ks >> 1) * a_step;
              const uint64_t b16 = b1_desc0 + b_lo_off + (uint64_t)(ks >> 1) * b1_step;
              mma_bf16<kPair>(tmem_d1, a16 + (a_lo_off >> 1), b16, idesc1h, 1);   // lo(a) . b
              mma_bf16<kPair>(tmem_d1, a16, b16 + (b_lo_off >> 1), idesc1h, 1);   // a . lo(b)
            }
            a_hi += a_step;
            b_hi += b1_step;
          }
          mma_commit_all<kPair>(&bars[BAR_D1_FULL]);
        }
        __syncwarp();
        // ---- layer 2: D2 = A2 * B2^T, A2 produced stage by stage (16 K-columns = two MMA K-steps) by epilogue-1 and
        //      handed over in TENSOR MEMORY (the MMA reads its A operand from there), B2 from shared memory.
        //      Software pipelined: the (100+ cycle) barrier probe for stage b+1 runs while the tensor pipe executes
        //      the 6 MMAs of stage b; the commit of stage b is issued after it (it still only covers MMAs <= b).
        const uint64_t b_lo_off = (uint64_t)(b2_hl_bytes >> 4);
        mbar_wait(&bars[BAR_A2_FULL + ring_st], ring_par);
        mbar_wait(&bars[BAR_D2_EMPTY], tpar ^ 1u);   // epilogue-2 of the previous tile drained D2
        tc::fence_after_thread_sync();
        for (int b = 0; b < n_b2; ++b) {
          trace_ev(trs, 0, tn, 2, tile_cnt * n_b2 + b);
          const uint32_t cur_st = ring_st;
          // Operands of this stage, computed HERE in convergent code so that they live in uniform registers: values
          // produced inside the single-lane block below reach the UTCHMMA through one R2UR each, and that serial
          // chain (not the tensor pipe) used to set the stage period.
          const uint32_t a0 = tmem_a2 + cur_st * (uint32_t)kTcStageCols;
          const uint64_t bd0 = b2_desc0 + (uint64_t)(2 * b) * b2_step, bd1 = bd0 + b2_step;
          const uint32_t acc0 = b > 0 ? 1u : 0u;
          const bool two = 2 * b + 1 < n_k2;   // the last stage of a 3xTF32 / TF32 tile may hold a single K-step
          // probe the next stage's barrier now: the probe's latency overlaps the MMA issue below, and in steady state
          // (epilogue-1 runs a stage or two ahead) it already succeeds, so the blocking wait after the issue is skipped
          uint32_t nxt_st = ring_st + 1, nxt_par = ring_par;
          if (nxt_st == (uint32_t)p.ring) { nxt_st = 0; nxt_par ^= 1u; }
          const bool nxt_ready = (b + 1 < n_b2) && mbar_test(&bars[BAR_A2_FULL + nxt_st], nxt_par);
          if (tc::elect_one()) {
            if (three) {
              mma_tf32_ts<kPair>(tmem_d2, a0, bd0, idesc2, acc0);
              mma_tf32_ts<kPair>(tmem_d2, a0 + 16u, bd0, idesc2, 1);
              mma_tf32_ts<kPair>(tmem_d2, a0, bd0 + b_lo_off, idesc2, 1);
              if (two) {
                mma_tf32_ts<kPair>(tmem_d2, a0 + 8u, bd1, idesc2, 1);
                mma_tf32_ts<kPair>(tmem_d2, a0 + 24u, bd1, idesc2, 1);
                mma_tf32_ts<kPair>(tmem_d2, a0 + 8u, bd1 + b_lo_off, idesc2, 1);
              }
            } else if (mixed) {
              // stage columns: tf32 hi [0,16) | bf16(a) pairs [16,24) | bf16(a_lo) pairs [24,32); the BF16 images of B2
              // sit behind its TF32 image, one K = 16 block per stage
              const uint64_t b16 = b2_desc0 + b_lo_off + (uint64_t)b * b2_step;
              mma_tf32_ts<kPair>(tmem_d2, a0, bd0, idesc2, acc0);
              mma_tf32_ts<kPair>(tmem_d2, a0 + 8u, bd1, idesc2, 1);
              mma_bf16_ts<kPair>(tmem_d2, a0 + 24u, b16, idesc2h, 1);                     // lo(a) . b
              mma_bf16_ts<kPair>(tmem_d2, a0 + 16u, b16 + (b_lo_off >> 1), idesc2h, 1);   // a . lo(b)
            } else {
              mma_tf32_ts<kPair>(tmem_d2, a0, bd0, idesc2, acc0);
              if (two) mma_tf32_ts<kPair>(tmem_d2, a0 + 8u, bd1, idesc2, 1);
            }
            // all MMAs issued so far complete -> the stage may be overwritten
            mma_commit_all<kPair>(&bars[BAR_A2_EMPTY + cur_st]);
          }
          __syncwarp();
          trace_ev(trs, 0, tn, 3, tile_cnt * n_b2 + b);
          ring_st = nxt_st; ring_par = nxt_par;
          if (b + 1 < n_b2) {
            trace_ev(trs, 0, tn, 1, tile_cnt * n_b2 + b + 1);
            if (!nxt_ready) mbar_wait(&bars[BAR_A2_FULL + ring_st], ring_par);
            tc::fence_after_thread_sync();
          }
        }
        if (tc::elect_one()) {
          mma_commit_all<kPair>(&bars[BAR_D2_FULL]);
          if (pt + 1 == pt1) mma_commit_all<kPair>(&bars[BAR_SAMPLE_DONE]);
        }
        __syncwarp();
        trace_ev(tr, 0, tn, 21, tile_cnt);
      }
    }
  } else if (warp < kTcE1Warps) {
    // ===================== WG-A: A1 + epilogue 1; thread = one point (TMEM lane) ===============================
    if constexpr (kTcSets == 3 && BNN_TC_REG_E1 < 80) setmaxnreg_dec<BNN_TC_REG_E1>();
    if constexpr (kTcSets == 3 && BNN_TC_REG_E1 > 80) setmaxnreg_inc<BNN_TC_REG_E1>();
    const int quarter = warp & 3, set = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    uint32_t tile_cnt = 0, tn = 0;
    unsigned long long* tr = (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && quarter == 0 && (p.trace_mask & 2)) ? p.trace : nullptr;
    const int who = 1 + set;
    uint32_t ring_gb = 0, ring_st = 0, ring_use = 0;   // last stage index handled by this warp, its slot and use parity
    const bool three = p.passes == 3, mixed = p.passes == 2;
    const float* b1s = reinterpret_cast<const float*>(smem + p.sm_vec);
    // this group's masks m0[Dp] (inputs) | m1[N1] (hidden 1), double-buffered by sample parity: with a shared network the
    // next sample's first tile is prepared while the current sample's last tile is still in flight
    float* mk_base = reinterpret_cast<float*>(smem + p.sm_mask);
    const int mk_sz = p.Dp + p.N1;
    float* mk0 = mk_base;
    float* mk1 = mk0 + p.Dp;
    const bool has_m0 = p.mask_mode != BNN_MASK_NONE && p.m_off[0] >= 0;
    const bool has_m1 = p.mask_mode != BNN_MASK_NONE && p.m_off[1] >= 0;
    // A1 granule kc of point n: hi/lo split of X[n][4kc .. 4kc+3] (zero beyond D / N)
    auto store_a1 = [&](int kc, const float (&x)[4]) {
      float4 hi, lo;
      hi.x = tc::tf32_hi(x[0]); hi.y = tc::tf32_hi(x[1]); hi.z = tc::tf32_hi(x[2]); hi.w = tc::tf32_hi(x[3]);
      lo.x = x[0] - hi.x; lo.y = x[1] - hi.y; lo.z = x[2] - hi.z; lo.w = x[3] - hi.w;
      const uint32_t a = s_a1 + (uint32_t)(kc * 128 + row) * 16u;
      sts128(a, hi);
      if (mixed) {   // BF16 granule = 8 k: this 4-k granule fills half of one
        const uint32_t a16 = s_a1 + a1_hl_bytes + (uint32_t)((kc >> 1) * 128 + row) * 16u + (uint32_t)(kc & 1) * 8u;
        sts64(a16, make_float2(__uint_as_float(tc::pack_bf16x2(x[0], x[1])), __uint_as_float(tc::pack_bf16x2(x[2], x[3]))));
        sts64(a16 + a1_hl_bytes / 2,
              make_float2(__uint_as_float(tc::pack_bf16x2(lo.x, lo.y)), __uint_as_float(tc::pack_bf16x2(lo.z, lo.w))));
      } else {
        sts128(a + a1_hl_bytes, lo);
      }
    };
    auto load_x = [&](long long n, int kc, float (&x)[4]) {
      const float* xr = p.X + n * p.D;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = kc * 4 + e;
        x[e] = (n < p.N && k < p.D) ? __ldg(xr + k) : 0.0f;
      }
    };
    auto mask_x = [&](const float* m0, int kc, float (&x)[4]) {   // dropout on the network inputs (BNN_Dropout.py:73-74 when dim > 1)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = kc * 4 + e;
        if (k < p.D) x[e] *= m0[k];
      }
    };
    constexpr int kPre = 4;   // granules of the next tile's input row prefetched into registers (D <= 16 fully)
    const int n_kc = p.Dp >> 2;
    auto draw_masks = [&](int sm) {   // masks of sample sm into the buffer of its parity (all 256 WG-A threads)
      float* m0 = mk_base + ((sm - s0) & 1) * mk_sz;
      float* m1 = m0 + p.Dp;
      const int t256 = warp * 32 + lane;
      if (has_m0) for (int j = t256; j < p.D; j += kTcE1Threads) m0[j] = tc_mask_value(p, p.sample_offset + sm, sm, 0, p.m_off[0] + j);
      if (has_m1) for (int j = t256; j < p.N1; j += kTcE1Threads) m1[j] = j < p.H1 ? tc_mask_value(p, p.sample_offset + sm, sm, 1, p.m_off[1] + j) : 0.0f;
    };
    for (int s = s0; s < s1; ++s) {
      const bool reload = !p.shared_w || s == s0;
      // With a shared network nothing is reloaded between samples, so (sample, tile) pairs form one continuous stream: the
      // pair after the last tile of sample s is the first tile of sample s + 1, prepared like any next tile.
      const bool chain = p.shared_w && s + 1 < s1;
      if (has_m0 || has_m1) {
        // masks are drawn ONE SAMPLE AHEAD (the next sample's first A1 needs them during this sample's last tile): every
        // WG-A warp has left sample s - 1, whose buffer is the one overwritten now (named barrier 5)
        named_bar_sync(5, kTcE1Threads);
        if (s == s0) draw_masks(s0);
        if (s + 1 < s1) draw_masks(s + 1);
        named_bar_sync(5, kTcE1Threads);
      }
      mk0 = mk_base + ((s - s0) & 1) * mk_sz;
      mk1 = mk0 + p.Dp;
      const float* mk0_next = mk_base + ((s + 1 - s0) & 1) * mk_sz;
      for (int pt = pt0; pt < pt1; ++pt, ++tile_cnt) {
        const uint32_t tpar = tile_cnt & 1u;
        float xn[kPre][4];
        const bool last_pt = pt + 1 == pt1;
        const bool have_next = !last_pt || chain;
        const int pt_next = last_pt ? pt0 : pt + 1;
        const float* m0_next = last_pt ? mk0_next : mk0;
        const long long n_next = ((long long)pt_next * kPair + rank) * 128 + row;
        const bool prepared = pt > pt0 || (p.shared_w && s > s0);   // A1 of this pair was built during the previous pair
        if (set == 0) {
          if (!prepared) {
            // first tile of the stream / of a freshly loaded sample: build A1 on the critical path; the operands of this
            // sample must have landed in THIS CTA before the leader may use them, so wait for them before signalling
            const long long n = ((long long)pt * kPair + rank) * 128 + row;
            for (int kc = 0; kc < n_kc; ++kc) {
              float x[4];
              load_x(n, kc, x);
              if (has_m0) mask_x(mk0, kc, x);
              store_a1(kc, x);
            }
            if (reload) mbar_wait(&bars[BAR_W_FULL], (uint32_t)(s - s0) & 1u);
            warp_signal_leader<kPair>(&bars[BAR_A1_FULL], lane);
          }
          if (have_next) {
#pragma unroll
            for (int kc = 0; kc < kPre; ++kc)
              if (kc < n_kc) load_x(n_next, kc, xn[kc]);
          }
        } else if (pt == pt0 && reload) {
          mbar_wait(&bars[BAR_W_FULL], (uint32_t)(s - s0) & 1u);   // b1 of this sample
        }

        mbar_wait(&bars[BAR_D1_FULL], tpar);
        tc::fence_after_thread_sync();
        // MMA1 of this tile has consumed A1: build the next tile's A1 now, so that MMA1(t+1) can run right behind
        // MMA2(t) (by then epilogue-1 of this tile has finished reading D1)
        if (set == 0 && have_next) {
#pragma unroll
          for (int kc = 0; kc < kPre; ++kc)
            if (kc < n_kc) {
              if (has_m0) mask_x(m0_next, kc, xn[kc]);
              store_a1(kc, xn[kc]);
            }
          for (int kc = kPre; kc < n_kc; ++kc) {
            float x[4];
            load_x(n_next, kc, x);
            if (has_m0) mask_x(m0_next, kc, x);
            store_a1(kc, x);
          }
          warp_signal_leader<kPair>(&bars[BAR_A1_FULL], lane);
        }

        // ---- epilogue 1: D1 -> h1 = act(d + b1) -> hi/lo -> A2 ring; one 16-column batch = one ring stage; this warp
        //      takes the batches whose index parity equals `set`
        // The D1 columns of this warp's NEXT stage are requested before the (long) wait for a free ring slot, so the
        // TMEM load latency never sits on the stage's critical path; masked / unmasked versions are separate code.
        auto epilogue1 = [&](auto masked_tag) {
          constexpr bool kMasked = decltype(masked_tag)::value;
          uint32_t d[16];
          if (set < n_b2) tc::tmem_ld16_issue(tmem_d1 + lane_sel + (uint32_t)(set * 16), d);
          for (int b = set; b < n_b2; b += kTcSets) {
            const int c0 = b * 16;
            const uint32_t gb = tile_cnt * (uint32_t)n_b2 + (uint32_t)b;   // global stage-use index
            // ring slot / use parity of stage gb, advanced incrementally (gb % ring and gb / ring cost ~25 instructions)
            ring_st += gb - ring_gb;
            ring_gb = gb;
            while (ring_st >= (uint32_t)p.ring) { ring_st -= (uint32_t)p.ring; ring_use ^= 1u; }
            trace_ev(tr, who, tn, 10, gb);
            tc::tmem_ld_wait(d);
            trace_ev(tr, who, tn, 11, gb);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 bq = *reinterpret_cast<const float4*>(b1s + c0 + q4 * 4);
              float h[4];
              h[0] = act_ct<kAct>(__uint_as_float(d[q4 * 4 + 0]) + bq.x);
              h[1] = act_ct<kAct>(__uint_as_float(d[q4 * 4 + 1]) + bq.y);
              h[2] = act_ct<kAct>(__uint_as_float(d[q4 * 4 + 2]) + bq.z);
              h[3] = act_ct<kAct>(__uint_as_float(d[q4 * 4 + 3]) + bq.w);
              if constexpr (kMasked) {   // dropout on the inputs of Linear 2
                const float4 mq = *reinterpret_cast<const float4*>(mk1 + c0 + q4 * 4);
                h[0] *= mq.x; h[1] *= mq.y; h[2] *= mq.z; h[3] *= mq.w;
              }
              float l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float hh = tc::tf32_hi_fast(h[e]);
                hi[q4 * 4 + e] = __float_as_uint(hh);
                l[e] = h[e] - hh;
              }
              if (mixed) {
                lo[q4 * 2 + 0] = tc::pack_bf16x2(h[0], h[1]);
                lo[q4 * 2 + 1] = tc::pack_bf16x2(h[2], h[3]);
                lo[8 + q4 * 2 + 0] = tc::pack_bf16x2(l[0], l[1]);
                lo[8 + q4 * 2 + 1] = tc::pack_bf16x2(l[2], l[3]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) lo[q4 * 4 + e] = __float_as_uint(l[e]);
              }
            }
            if (b + kTcSets < n_b2) tc::tmem_ld16_issue(tmem_d1 + lane_sel + (uint32_t)(c0 + 16 * kTcSets), d);   // next stage of this warp
            const uint32_t st = ring_st;
            trace_ev(tr, who, tn, 12, gb);
            mbar_wait(&bars[BAR_A2_EMPTY + st], ring_use ^ 1u);   // the MMAs that read this stage's previous contents are done
            tc::fence_after_thread_sync();
            trace_ev(tr, who, tn, 13, gb);
            const uint32_t ta = tmem_a2 + lane_sel + st * (uint32_t)kTcStageCols;
            tc::tmem_st16(ta, hi);
            if (three || mixed) tc::tmem_st16(ta + 16u, lo);   // tf32 lo [16,32)  or  bf16(h) pairs [16,24) | bf16(lo) pairs [24,32)
            tc::tmem_st_wait();
            warp_signal_tmem<kPair>(&bars[BAR_A2_FULL + st], lane);
            trace_ev(tr, who, tn, 14, gb);
          }
        };
        if (has_m1) epilogue1(std::true_type{});
        else epilogue1(std::false_type{});
      }
    }
  } else if (warp < kTcMmaWarp) {
    // ===================== WG-B: epilogue 2 + predictive reduction; thread = one point ===============================
    if constexpr (kTcSets == 3 && BNN_TC_REG_E2 > 80) setmaxnreg_inc<BNN_TC_REG_E2>();
    const int quarter = warp & 3, half = (warp - kTcE1Warps) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    uint32_t tile_cnt = 0, tn = 0;
    unsigned long long* tr = (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && quarter == 0 && (p.trace_mask & 4)) ? p.trace : nullptr;
    const int who = 1 + kTcSets + half;
    const long long M = p.N * p.O;
    double* st_mean = p.part ? p.part + (long long)g * 2 * M : nullptr;
    double* st_m2 = p.part ? st_mean + M : nullptr;
    const float* b3 = reinterpret_cast<const float*>(smem + p.sm_vec) + p.N1 + p.N2 + p.O * p.N2;
    float* ypart = reinterpret_cast<float*>(smem + p.sm_ypart);   // [128][O] partial outputs of the half-1 warps
    const int nb16 = p.N2 >> 4;
    const int c_lo = half == 0 ? 0 : ((nb16 + 1) >> 1) * 16, c_hi = half == 0 ? ((nb16 + 1) >> 1) * 16 : p.N2;
    // with dropout on the inputs of the last Linear, w3 is replaced by a per-sample copy with its columns scaled by the
    // mask -- exactly what the reference does to the weight matrix (BNN_Dropout.py:70-75, BNN_CDropout.py:59-65)
    float* w3m = reinterpret_cast<float*>(smem + p.sm_mask) + 2 * (p.Dp + p.N1);   // [O][N2]
    const float* b2s = reinterpret_cast<const float*>(smem + p.sm_vec) + p.N1;
    const float* w3s = b2s + p.N2;
    const bool has_m2 = p.mask_mode != BNN_MASK_NONE && p.m_off[2] >= 0;
    for (int s = s0; s < s1; ++s) {
      const double inv_cnt = 1.0 / (double)(s - s0 + 1);
      if (!p.shared_w || s == s0) mbar_wait(&bars[BAR_W_FULL], (uint32_t)(s - s0) & 1u);
      if (has_m2) {
        named_bar_sync(6, 256);
        const int t256 = (warp - kTcE1Warps) * 32 + lane;
        for (int j = t256; j < p.N2; j += 256) {
          const float m = j < p.H2 ? tc_mask_value(p, p.sample_offset + s, s, 2, p.m_off[2] + j) : 0.0f;
          for (int o = 0; o < p.O; ++o) w3m[o * p.N2 + j] = w3s[o * p.N2 + j] * m;
        }
        named_bar_sync(6, 256);
      }
      float y0[4] = {0.f, 0.f, 0.f, 0.f};
      if (half == 0) { y0[0] = b3[0]; if (p.O > 1) y0[1] = b3[1]; if (p.O > 2) y0[2] = b3[2]; if (p.O > 3) y0[3] = b3[3]; }
      for (int pt = pt0; pt < pt1; ++pt, ++tile_cnt) {
        const uint32_t tpar = tile_cnt & 1u;
        const long long n = ((long long)pt * kPair + rank) * 128 + row;
        const bool in_range = n < p.N;
        // ---- epilogue 2: D2 -> y[o] = b3[o] + sum_j w3[o][j] * act(d + b2[j]) over this warp's column half; the TMEM
        //      load of batch i+1 is in flight while batch i is reduced (4 independent partial sums per output)
        trace_ev(tr, who, tn, 30, tile_cnt);
        mbar_wait(&bars[BAR_D2_FULL], tpar);
        tc::fence_after_thread_sync();
        trace_ev(tr, who, tn, 31, tile_cnt);
        float y[4] = {y0[0], y0[1], y0[2], y0[3]};
        // Every tcgen05.ld / wait::ld round trip costs a few hundred cycles here regardless of its width (8 warps and the
        // next tile's layer-1 MMAs share the TMEM ports), so D2 is drained with 32-column loads and the
        // activations are consumed four columns at a time to keep the register count under the 120 the CTA allows.
        // kSingle (one output: every configuration the reference ships but the multi-output toy) keeps four running
        // partial sums in registers; the general case predicates an unrolled loop so y[] is never indexed at run time.
        auto drain = [&](auto single_tag) {
          constexpr bool kSingle = decltype(single_tag)::value;
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          const float* w3p = has_m2 ? w3m : w3s;
          // the 16-column constants are plain shared-memory loads the compiler may schedule freely (the asm wrappers
          // serialised one exposed LDS latency per four columns); the first set is fetched before the TMEM wait
          float4 bq[4], wq[4];
          auto fetch = [&](int c) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              bq[q4] = *reinterpret_cast<const float4*>(b2s + c + q4 * 4);
              if constexpr (kSingle) wq[q4] = *reinterpret_cast<const float4*>(w3p + c + q4 * 4);
            }
          };
          auto reduce16 = [&](const uint32_t* d, int c) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float t0 = act_ct<kAct>(__uint_as_float(d[q4 * 4 + 0]) + bq[q4].x);
              const float t1 = act_ct<kAct>(__uint_as_float(d[q4 * 4 + 1]) + bq[q4].y);
              const float t2 = act_ct<kAct>(__uint_as_float(d[q4 * 4 + 2]) + bq[q4].z);
              const float t3 = act_ct<kAct>(__uint_as_float(d[q4 * 4 + 3]) + bq[q4].w);
              if constexpr (kSingle) {
                acc[0] = fmaf(wq[q4].x, t0, acc[0]);
                acc[1] = fmaf(wq[q4].y, t1, acc[1]);
                acc[2] = fmaf(wq[q4].z, t2, acc[2]);
                acc[3] = fmaf(wq[q4].w, t3, acc[3]);
              } else {
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                  if (o < p.O) {
                    const float4 w4 = *reinterpret_cast<const float4*>(w3p + o * p.N2 + c + q4 * 4);
                    acc[o] += fmaf(w4.x, t0, w4.y * t1) + fmaf(w4.z, t2, w4.w * t3);
                  }
                }
              }
            }
          };
          uint32_t d[32];
          for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            const bool wide = c_hi - c0 >= 32;
            if (wide) tc::tmem_ld32_issue(tmem_d2 + lane_sel + (uint32_t)c0, d);
            else tc::tmem_ld16_issue_lo(tmem_d2 + lane_sel + (uint32_t)c0, d);
            fetch(c0);
            tc::tmem_ld_wait32(d);
            reduce16(d, c0);
            if (wide) {
              fetch(c0 + 16);
              reduce16(d + 16, c0 + 16);
            }
          }
          if constexpr (kSingle) y[0] += (acc[0] + acc[1]) + (acc[2] + acc[3]);
          else {
#pragma unroll
            for (int o = 0; o < 4; ++o) y[o] += acc[o];
          }
        };
        if (p.O == 1) drain(std::true_type{});
        else drain(std::false_type{});
        trace_ev(tr, who, tn, 32, tile_cnt);
        // D2 is drained by this warp: let the next tile's layer-2 MMAs overwrite it; after the last tile of the sample
        // the bias vector is free as well
        tc::fence_before_thread_sync();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kPair == 1) mbar_arrive(&bars[BAR_D2_EMPTY]);
          else mbar_arrive_cluster(&bars[BAR_D2_EMPTY], 0);
          if (pt + 1 == pt1) mbar_arrive(&bars[BAR_VEC_FREE]);
        }
        // this point's Welford state: fetched only now, off the critical path (D2 is already released and this warpgroup
        // idles until the next tile's layer 2 is done), so that it holds no registers while D2 is drained
        double mu0[4] = {0.0, 0.0, 0.0, 0.0}, q0[4] = {0.0, 0.0, 0.0, 0.0};
        if (half == 0 && in_range && st_mean && s != s0) {
#pragma unroll
          for (int o = 0; o < 4; ++o) {   // unrolled + predicated: mu0/q0/y stay in registers
            if (o < p.O) {
              mu0[o] = st_mean[n * p.O + o];
              q0[o] = st_m2[n * p.O + o];
            }
          }
        }
        // the two column halves of a lane quarter meet in shared memory (named barrier 1 + quarter, 64 threads)
        if (half == 1) {
#pragma unroll
          for (int o = 0; o < 4; ++o) if (o < p.O) ypart[row * p.O + o] = y[o];
        }
        named_bar_sync(1 + quarter, 64);
        if (half == 0) {
#pragma unroll
          for (int o = 0; o < 4; ++o) if (o < p.O) y[o] += ypart[row * p.O + o];
        }
        named_bar_sync(1 + quarter, 64);   // ypart may be overwritten by the next tile
        // ---- predictive reduction: FP64 Welford state of this point (only this thread ever touches it)
        if (half == 0 && in_range) {
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            if (o < p.O) {
              const long long m = n * p.O + o;
              if (p.pred) p.pred[((long long)s * p.N + n) * p.O + o] = y[o];
              if (st_mean) {
                const double dlt = (double)y[o] - mu0[o];
                const double mu = mu0[o] + dlt * inv_cnt;
                st_mean[m] = mu;
                st_m2[m] = q0[o] + dlt * ((double)y[o] - mu);
              }
            }
          }
        }
      }
    }
  }

  // ---- teardown
  tc::fence_before_thread_sync();
  __syncthreads();
  if constexpr (kPair == 2) cluster_sync_all();
  if (warp == kTcMmaWarp) tmem_dealloc_n<kPair>(tmem_base, (uint32_t)p.tmem_cols);
}

// ---- self-test GEMM: D[128, N] = A[128, K] * B[N, K]^T through the same descriptors / split -----------------
__global__ void __launch_bounds__(160, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             float* __restrict__ Dout, int K, int N, int passes) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_hl = (uint32_t)K * 128 * 4, b_hl = (uint32_t)K * N * 4;
  const uint32_t s_a = smem_u32(smem), s_b = s_a + 2 * a_hl;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  const bool a_in_tmem = (passes & 16) != 0;   // A operand staged in tensor memory (columns 256.. : hi[K], lo[K]; K <= 128)
  passes &= 15;
  if (warp == 4) tc::tmem_alloc(&tmem_slot, 512);
  if (passes == 2) {
    // ---- mixed scheme: a.b ~ tf32(a).tf32(b) [kind::tf32] + bf16(a - tf32(a)).bf16(b) + bf16(a).bf16(b - tf32(b)) [kind::f16];
    //      per operand: tf32 image (16-byte granule = 4 k), then two bf16 images (16-byte granule = 8 k) in the second half
    auto fill = [&](const float* src, int rows, uint32_t s_dst, uint32_t hl) {
      for (int i = tid; i < (K / 8) * rows; i += blockDim.x) {
        const int kc8 = i / rows, r = i % rows;
        float x[8], h[8];
        for (int e = 0; e < 8; ++e) { x[e] = src[(size_t)r * K + kc8 * 8 + e]; h[e] = tc::tf32_hi(x[e]); }
        sts128(s_dst + (uint32_t)((2 * kc8) * rows + r) * 16u, make_float4(h[0], h[1], h[2], h[3]));
        sts128(s_dst + (uint32_t)((2 * kc8 + 1) * rows + r) * 16u, make_float4(h[4], h[5], h[6], h[7]));
        uint4 ph, pl;
        ph.x = tc::pack_bf16x2(x[0], x[1]); ph.y = tc::pack_bf16x2(x[2], x[3]);
        ph.z = tc::pack_bf16x2(x[4], x[5]); ph.w = tc::pack_bf16x2(x[6], x[7]);
        pl.x = tc::pack_bf16x2(x[0] - h[0], x[1] - h[1]); pl.y = tc::pack_bf16x2(x[2] - h[2], x[3] - h[3]);
        pl.z = tc::pack_bf16x2(x[4] - h[4], x[5] - h[5]); pl.w = tc::pack_bf16x2(x[6] - h[6], x[7] - h[7]);
        *reinterpret_cast<uint4*>(smem + (s_dst - s_a) + hl + (size_t)i * 16) = ph;
        *reinterpret_cast<uint4*>(smem + (s_dst - s_a) + hl + hl / 2 + (size_t)i * 16) = pl;
      }
    };
    fill(A, 128, s_a, a_hl);
    fill(B, N, s_b, b_hl);
    fence_proxy_async();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    const uint32_t tmem = tmem_slot;
    if (a_in_tmem && warp < 4) {   // tensor memory: hi[K] | bf16 hi pairs [K/2] | bf16 lo pairs [K/2] from column 256
      const int r = warp * 32 + lane;
      const uint32_t base = tmem + ((uint32_t)(warp * 32) << 16) + 256u;
      for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t hi[16], ph[8], pl[8];
        float x[16], h[16];
        for (int j = 0; j < 16; ++j) { x[j] = A[(size_t)r * K + k0 + j]; h[j] = tc::tf32_hi(x[j]); hi[j] = __float_as_uint(h[j]); }
        for (int j = 0; j < 8; ++j) {
          ph[j] = tc::pack_bf16x2(x[2 * j], x[2 * j + 1]);
          pl[j] = tc::pack_bf16x2(x[2 * j] - h[2 * j], x[2 * j + 1] - h[2 * j + 1]);
        }
        tc::tmem_st16(base + (uint32_t)k0, hi);
        tc::tmem_st8(base + (uint32_t)(K + k0 / 2), ph);
        tc::tmem_st8(base + (uint32_t)(K + K / 2 + k0 / 2), pl);
      }
      tc::tmem_st_wait();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
    if (warp == 4 && lane == 0) {
      const uint32_t id32 = tc::make_idesc_tf32(128, N), id16 = tc::make_idesc_bf16(128, N);
      for (int k16 = 0; k16 < K / 16; ++k16) {
        const uint64_t b16h = tc::make_smem_desc(s_b + b_hl + (uint32_t)(2 * k16) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
        const uint64_t b16l = tc::make_smem_desc(s_b + b_hl + b_hl / 2 + (uint32_t)(2 * k16) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
        for (int j = 0; j < 2; ++j) {
          const int ks = 2 * k16 + j;
          const uint64_t bh = tc::make_smem_desc(s_b + (uint32_t)(2 * ks) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
          if (a_in_tmem) tc::mma_tf32_ts(tmem, tmem + 256u + (uint32_t)(ks * 8), bh, id32, ks > 0);
          else tc::mma_tf32_ss(tmem, tc::make_smem_desc(s_a + (uint32_t)(2 * ks) * 2048u, 2048u, 128u), bh, id32, ks > 0);
        }
        if (a_in_tmem) {
          tc::mma_bf16_ts(tmem, tmem + 256u + (uint32_t)(K + K / 2 + k16 * 8), b16h, id16, 1);   // lo(a) . b
          tc::mma_bf16_ts(tmem, tmem + 256u + (uint32_t)(K + k16 * 8), b16l, id16, 1);           // a . lo(b)
        } else {
          tc::mma_bf16_ss(tmem, tc::make_smem_desc(s_a + a_hl + a_hl / 2 + (uint32_t)(2 * k16) * 2048u, 2048u, 128u), b16h, id16, 1);
          tc::mma_bf16_ss(tmem, tc::make_smem_desc(s_a + a_hl + (uint32_t)(2 * k16) * 2048u, 2048u, 128u), b16l, id16, 1);
        }
      }
      tc::mma_commit(&bar);
    }
    if (warp < 4) {
      mbar_wait(&bar, 0);
      tc::fence_after_thread_sync();
      const int row = warp * 32 + lane;
      for (int c0 = 0; c0 < N; c0 += 16) {
        float d[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, d);
        for (int j = 0; j < 16; ++j) Dout[(size_t)row * N + c0 + j] = d[j];
      }
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem, 512);
    return;
  }
  // operands -> K-major granules, hi/lo
  for (int i = tid; i < (K / 4) * 128; i += blockDim.x) {
    const int kc = i / 128, r = i % 128;
    float4 x = *reinterpret_cast<const float4*>(A + (size_t)r * K + kc * 4);
    float4 hi = make_float4(tc::tf32_hi(x.x), tc::tf32_hi(x.y), tc::tf32_hi(x.z), tc::tf32_hi(x.w));
    float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
    sts128(s_a + (uint32_t)i * 16u, hi);
    sts128(s_a + a_hl + (uint32_t)i * 16u, lo);
  }
  for (int i = tid; i < (K / 4) * N; i += blockDim.x) {
    const int kc = i / N, r = i % N;
    float4 x = *reinterpret_cast<const float4*>(B + (size_t)r * K + kc * 4);
    float4 hi = make_float4(tc::tf32_hi(x.x), tc::tf32_hi(x.y), tc::tf32_hi(x.z), tc::tf32_hi(x.w));
    float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
    sts128(s_b + (uint32_t)i * 16u, hi);
    sts128(s_b + b_hl + (uint32_t)i * 16u, lo);
  }
  fence_proxy_async();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_slot;
  if (a_in_tmem) {
    if (warp < 4) {
      const int r = warp * 32 + lane;
      const uint32_t base = tmem + ((uint32_t)(warp * 32) << 16) + 256u;
      for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t hi[16], lo[16];
        for (int j = 0; j < 16; ++j) {
          const float x = k0 + j < K ? A[(size_t)r * K + k0 + j] : 0.0f;
          const float h = tc::tf32_hi(x);
          hi[j] = __float_as_uint(h);
          lo[j] = __float_as_uint(x - h);
        }
        tc::tmem_st16(base + (uint32_t)k0, hi);
        tc::tmem_st16(base + (uint32_t)(K + k0), lo);
      }
      tc::tmem_st_wait();
    }
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
  }
  if (warp == 4 && lane == 0 && a_in_tmem) {
    const uint32_t idesc = tc::make_idesc_tf32(128, N);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t a_hi = tmem + 256u + (uint32_t)(ks * 8), a_lo = a_hi + (uint32_t)K;
      const uint64_t b_hi = tc::make_smem_desc(s_b + (uint32_t)(2 * ks) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
      const uint64_t b_lo = tc::make_smem_desc(s_b + b_hl + (uint32_t)(2 * ks) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
      tc::mma_tf32_ts(tmem, a_hi, b_hi, idesc, ks > 0);
      if (passes == 3) {
        tc::mma_tf32_ts(tmem, a_lo, b_hi, idesc, 1);
        tc::mma_tf32_ts(tmem, a_hi, b_lo, idesc, 1);
      }
    }
    tc::mma_commit(&bar);
  } else if (warp == 4 && lane == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(128, N);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t a_hi = tc::make_smem_desc(s_a + (uint32_t)(2 * ks) * 2048u, 2048u, 128u);
      const uint64_t a_lo = tc::make_smem_desc(s_a + a_hl + (uint32_t)(2 * ks) * 2048u, 2048u, 128u);
      const uint64_t b_hi = tc::make_smem_desc(s_b + (uint32_t)(2 * ks) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
      const uint64_t b_lo = tc::make_smem_desc(s_b + b_hl + (uint32_t)(2 * ks) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
      tc::mma_tf32_ss(tmem, a_hi, b_hi, idesc, ks > 0);
      if (passes == 3) {
        tc::mma_tf32_ss(tmem, a_lo, b_hi, idesc, 1);
        tc::mma_tf32_ss(tmem, a_hi, b_lo, idesc, 1);
      }
    }
    tc::mma_commit(&bar);
  }
  if (warp < 4) {
    mbar_wait(&bar, 0);
    tc::fence_after_thread_sync();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float d[16];
      tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, d);
      for (int j = 0; j < 16; ++j) Dout[(size_t)row * N + c0 + j] = d[j];
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

// ---- MMA issue-rate probe (debug): `iters` back-to-back MMAs of one CTA (M = 128) into one accumulator, timed with
//      clock64 from the first issue to the completion barrier.  mode: 0 tf32 SS, 1 tf32 TS, 2 bf16 SS, 3 bf16 TS,
//      4 = the mixed stage pattern (tf32 TS x2 + bf16 TS x2), 5 = the 3xTF32 stage pattern (tf32 TS x6).
//      Operand contents are whatever shared / tensor memory holds: only the timing is meaningful.
__global__ void __launch_bounds__(160, 1) tc_mma_bench_kernel(int mode, int N, int iters, long long* out_clk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 4) tc::tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 4) {   // converged warp, one elected lane issues (see the note in fwd_tc_kernel)
    const uint32_t s0 = smem_u32(smem);
    const uint64_t adesc = tc::make_smem_desc(s0, 2048u, 128u);
    const uint64_t bdesc = tc::make_smem_desc(s0 + 8192u, (uint32_t)N * 16u, 128u);
    const uint32_t id32 = tc::make_idesc_tf32(128, N), id16 = tc::make_idesc_bf16(128, N);
    const uint32_t ta = tmem + 256u;
    const long long t0 = clock64();
    if (tc::elect_one()) {
      if (mode == 0) for (int i = 0; i < iters; ++i) tc::mma_tf32_ss(tmem, adesc, bdesc, id32, 1);
      else if (mode == 1) for (int i = 0; i < iters; ++i) tc::mma_tf32_ts(tmem, ta, bdesc, id32, 1);
      else if (mode == 2) for (int i = 0; i < iters; ++i) tc::mma_bf16_ss(tmem, adesc, bdesc, id16, 1);
      else if (mode == 3) for (int i = 0; i < iters; ++i) tc::mma_bf16_ts(tmem, ta, bdesc, id16, 1);
      else if (mode == 4) {
        for (int i = 0; i < iters; i += 4) {
          tc::mma_tf32_ts(tmem, ta, bdesc, id32, 1);
          tc::mma_tf32_ts(tmem, ta + 8u, bdesc, id32, 1);
          tc::mma_bf16_ts(tmem, ta + 16u, bdesc, id16, 1);
          tc::mma_bf16_ts(tmem, ta + 24u, bdesc, id16, 1);
        }
      } else {
        for (int i = 0; i < iters; i += 3) {
          tc::mma_tf32_ts(tmem, ta, bdesc, id32, 1);
          tc::mma_tf32_ts(tmem, ta + 16u, bdesc, id32, 1);
          tc::mma_tf32_ts(tmem, ta, bdesc, id32, 1);
        }
      }
      tc::mma_commit(&bar);
    }
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (lane == 0) {
      out_clk[0] = t1 - t0;   // issue time
      out_clk[1] = t2 - t0;   // until all MMAs completed
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

// ---- host side ----------------------------------------------------------------------------------------------
static inline int roundup(int v, int m) { return (v + m - 1) / m * m; }

struct TcPlan {
  TcParams p;
  TcPrepParams q;
  int kpair;
  size_t smem;
  long long img_bytes, part_bytes, x_bytes, prog_bytes;   // x_bytes: copy of X next to the Welford state (one L2 window over both), or 0
  int slots;
};

int finalize_moments(const double* part, const int*, int G, int Sg, long long S, long long M, float* mean, float* var,
                     double* part_mean, double* part_m2, cudaStream_t st);

static long long l2_persist_cap();

// 0 = ok, 1 = shape not supported by the tensor-core path (message set), <0 = error
static int plan_tc(const BnnForwardArgs* a, TcPlan* plan) {
  Layout L;
  if (make_layout(&a->desc, &L)) return -1;
  if (L.n_lin != 3) { set_error("tensor-core forward needs exactly 2 hidden layers (got %d Linear layers)", L.n_lin); return 1; }
  const int D = L.in[0], H1 = L.out[0], H2 = L.out[1], O = L.out[2];
  if (D > 64 || H1 > 256 || H2 > 256 || O > 4) { set_error("tensor-core forward: shape outside D<=64, H<=256, O<=4"); return 1; }
  TcParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.D = D; p.H1 = H1; p.H2 = H2; p.O = O; p.act = L.act;
  p.passes = a->precision == BNN_PREC_TF32 ? 1 : a->precision == BNN_PREC_TF32_BF16 ? 2 : 3;
  const int kq = p.passes == 2 ? 16 : 8;   // the BF16 correction MMAs take K = 16 at a time
  p.Dp = roundup(D, kq); p.N1 = roundup(H1, 16); p.N2 = roundup(H2, 16); p.K2 = roundup(H1, kq);
  p.vec_len = p.N1 + p.N2 + O * p.N2 + 4;
  const int vec_pad = (p.vec_len + 3) & ~3;
  // try one CTA per tile first (weights fully resident in one SM's shared memory), else a CTA pair
  int kpair = 0;
  for (int kp = 1; kp <= 2 && !kpair; ++kp) {
    const int N1h = p.N1 / kp, N2h = p.N2 / kp;
    if ((N1h & 7) || (N2h & 7)) continue;
    {
      size_t o = 0;
      p.sm_b1 = (int)o; o += (size_t)2 * p.Dp * N1h * 4;
      p.sm_b2 = (int)o; o += (size_t)2 * p.K2 * N2h * 4;
      p.sm_a1 = (int)o; o += (size_t)2 * p.Dp * 128 * 4;
      p.sm_vec = (int)o; o += (size_t)vec_pad * 4;
      p.sm_ypart = (int)o; o += (size_t)128 * O * 4;
      o = (o + 15) & ~(size_t)15;
      p.sm_mask = (int)o; if (a->mask.mode != BNN_MASK_NONE) o += (size_t)(2 * (p.Dp + p.N1) + p.O * p.N2) * 4;
      o = (o + 15) & ~(size_t)15;
      p.sm_bar = (int)o; o += (size_t)BAR_COUNT * 8 + 16;
      if (o <= (size_t)227 * 1024) { kpair = kp; plan->smem = o; }
    }
  }
  if (!kpair) { set_error("tensor-core forward: weights do not fit shared memory even as a CTA pair"); return 1; }
  plan->kpair = kpair;
  const int N1h = p.N1 / kpair, N2h = p.N2 / kpair;
  p.part_stride = 2LL * p.Dp * N1h + 2LL * p.K2 * N2h;
  p.img_stride = p.part_stride * kpair + vec_pad;
  // tensor memory: D1 [N1] | D2 [N2] | A2 ring (as many 32-column stages as fit, 2..kTcRing)
  int ring = (512 - p.N1 - p.N2) / kTcStageCols;
  if (ring > kTcRing) ring = kTcRing;
  if (const char* e = getenv("BNN_TC_RING")) { const int v = atoi(e); if (v >= 2 && v < ring) ring = v; }  // tuning knob
  if (ring < 2) { set_error("tensor-core forward: accumulators and the activation ring exceed TMEM"); return 1; }
  p.ring = ring;
  int cols = p.N1 + p.N2 + ring * kTcStageCols, tc = 32;
  while (tc < cols) tc <<= 1;
  p.tmem_cols = tc;
  p.X = a->X; p.N = a->N; p.S = (int)a->S; p.pred = a->pred;
  p.shared_w = a->w_stride == 0;
  p.mask_mode = a->mask.mode;
  if (a->mask.mode != BNN_MASK_NONE) {
    int mo[BNN_MAX_LINEAR];
    p.mask_len = make_mask_offsets(L, a->mask.layer_masked, mo);
    for (int l = 0; l < 3; ++l) { p.m_off[l] = mo[l]; p.p_drop[l] = a->mask.p_drop[l]; }
    p.mask = a->mask.mask; p.mask_stride = a->mask.mask_stride; p.temperature = a->mask.temperature;
    p.seed = a->mask.seed; p.sample_offset = a->mask.sample_offset;
  } else {
    p.m_off[0] = p.m_off[1] = p.m_off[2] = -1;
  }
  // work split: slots over pair-tiles, sample groups when there are too few tiles
  const int tile_pts = 128 * kpair;
  p.n_tiles = (int)((a->N + tile_pts - 1) / tile_pts);
  const int max_slots = sm_count() / kpair;
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
  plan->img_bytes = ((long long)(p.shared_w ? 1 : a->S) * p.img_stride * 4 + 255) & ~255LL;
  plan->part_bytes = (long long)G * 2 * a->N * O * (long long)sizeof(double);
  // x_bytes stays 0: pinning a copy of X together with the Welford state (one 80 MB window at cfg 4) was measured WORSE than
  // pinning the state alone (20.7 vs 10.2 GB of DRAM traffic per 2048 samples; 35 GB with no window): the set-aside overflows.
  plan->x_bytes = 0;
  plan->prog_bytes = ((long long)G * max_slots * 4 + 255) & ~255LL;
  TcPrepParams& q = plan->q;
  q.D = D; q.H1 = H1; q.H2 = H2; q.O = O; q.Dp = p.Dp; q.N1 = p.N1; q.N2 = p.N2; q.K2 = p.K2; q.kpair = kpair;
  q.ld0 = L.ld[0]; q.ld1 = L.ld[1]; q.ld2 = L.ld[2];
  q.off0 = L.off[0]; q.off1 = L.off[1]; q.off2 = L.off[2];
  q.w_stride = a->w_stride; q.part_stride = p.part_stride; q.img_stride = p.img_stride; q.vec_len = p.vec_len; q.mixed = p.passes == 2;
  return 0;
}

// Bytes of the L2 set-aside a launch may pin (0: none).  The set-aside is sized once per device, on first use, to what the
// device allows (it stays usable by normal accesses whenever no persisting lines occupy it).  BNN_TC_L2PERSIST=0 turns it off.
static long long l2_persist_cap() {
  static long long cap[64];
  static bool done[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 0; }
  if (!done[dev]) {
    done[dev] = true;
    cap[dev] = 0;
    const char* e = getenv("BNN_TC_L2PERSIST");
    if (!(e && atoi(e) == 0)) {
      int max_persist = 0, max_win = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
      cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
      long long c = max_persist < max_win ? max_persist : max_win;
      if (c > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)c) == cudaSuccess) cap[dev] = c;
      cudaGetLastError();
    }
  }
  return cap[dev];
}

static unsigned long long* g_tc_trace = nullptr;
unsigned long long* tc_trace_buffer() { return g_tc_trace; }

long long forward_tc_workspace_bytes(const BnnForwardArgs* a) {
  TcPlan plan;
  const int rc = plan_tc(a, &plan);
  if (rc) return -1;
  return plan.img_bytes + plan.part_bytes + plan.x_bytes + plan.prog_bytes;
}

int forward_tc_supported(const BnnForwardArgs* a) {
  TcPlan plan;
  return plan_tc(a, &plan) == 0;
}

int forward_tc(const BnnForwardArgs* a, cudaStream_t st) {
  TcPlan plan;
  const int rc = plan_tc(a, &plan);
  if (rc) return -1;
  TcParams& p = plan.p;
  const long long need = plan.img_bytes + plan.part_bytes + plan.x_bytes + plan.prog_bytes;
  if (!a->workspace || a->workspace_bytes < need)
    BNN_FAIL("workspace too small: need %lld bytes, got %lld", need, (long long)a->workspace_bytes);
  float* img = reinterpret_cast<float*>(a->workspace);
  p.img = img;
  p.trace = nullptr;
#ifdef BNN_TC_TRACE_BUILD
  if (getenv("BNN_TC_TRACE")) {   // debug only: leaks one small buffer per call
    unsigned long long* tb = nullptr;
    if (cudaMalloc(&tb, 6 * 4096 * 16) == cudaSuccess) { cudaMemsetAsync(tb, 0, 6 * 4096 * 16, st); p.trace = tb; g_tc_trace = tb; p.trace_mask = atoi(getenv("BNN_TC_TRACE")); }
  }
#endif
  {
    p.progress = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(a->workspace) + plan.img_bytes + plan.part_bytes + plan.x_bytes);
    p.pace = 4;
    if (const char* e = getenv("BNN_TC_PACE")) p.pace = atoi(e);     // tuning knob; 0 = off
    if (p.pace <= 0 || p.shared_w) p.progress = nullptr;
    else BNN_CUDA(cudaMemsetAsync(p.progress, 0, (size_t)plan.prog_bytes, st));
  }
  const bool want_moments = a->mean || a->var || a->part_mean || a->part_m2;
  p.part = want_moments ? reinterpret_cast<double*>(reinterpret_cast<char*>(a->workspace) + plan.img_bytes) : nullptr;
  {
    long long bx = (p.img_stride + 255) / 256;
    if (bx > 64) bx = 64;
    plan.q.S = p.shared_w ? 1 : a->S;
    dim3 grid((unsigned)bx, (unsigned)(plan.q.S > 65535 ? 65535 : plan.q.S));
    tc_prep_kernel<<<grid, 256, 0, st>>>(plan.q, a->W, img);
    BNN_CUDA(cudaGetLastError());
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(plan.slots * plan.kpair), (unsigned)p.G, 1);
  cfg.blockDim = dim3(kTcThreads, 1, 1);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)plan.kpair;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // Pin the kernel's re-used working set -- the running (mean, M2) state and, when it fits, the copy of X behind it -- in the
  // L2 set-aside (see plan_tc).
  if (want_moments && plan.part_bytes > 0) {
    const long long cap = l2_persist_cap();
    long long win = plan.part_bytes + plan.x_bytes;
    if (win > cap) win = cap;
    if (win > 0) {
      if (plan.x_bytes > 0) {
        float* xc = reinterpret_cast<float*>(reinterpret_cast<char*>(a->workspace) + plan.img_bytes + plan.part_bytes);
        BNN_CUDA(cudaMemcpyAsync(xc, a->X, (size_t)a->N * p.D * 4, cudaMemcpyDeviceToDevice, st));
        p.X = xc;
      }
      attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
      attr[1].val.accessPolicyWindow.base_ptr = p.part;
      attr[1].val.accessPolicyWindow.num_bytes = (size_t)win;
      attr[1].val.accessPolicyWindow.hitRatio = 1.0f;
      attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
      cfg.numAttrs = 2;
    }
  }
  auto launch = [&](auto kern) -> int {
    BNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    BNN_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return 0;
  };
  int lrc = 0;
#define BNN_TC_LAUNCH(ACT)                                                         \
  case ACT: lrc = plan.kpair == 1 ? launch(fwd_tc_kernel<1, ACT>) : launch(fwd_tc_kernel<2, ACT>); break;
  switch (p.act) {
    BNN_TC_LAUNCH(BNN_ACT_RELU)
    BNN_TC_LAUNCH(BNN_ACT_TANH)
    BNN_TC_LAUNCH(BNN_ACT_SIGMOID)
    default: lrc = plan.kpair == 1 ? launch(fwd_tc_kernel<1, BNN_ACT_IDENTITY>) : launch(fwd_tc_kernel<2, BNN_ACT_IDENTITY>); break;
  }
#undef BNN_TC_LAUNCH
  if (lrc) return lrc;
  BNN_CUDA(cudaGetLastError());
  if (want_moments)
    return finalize_moments(p.part, nullptr, p.G, p.Sg, a->S, a->N * p.O, a->mean, a->var, a->part_mean, a->part_m2, st);
  return 0;
}

}  // namespace bnn

using namespace bnn;

namespace bnn { unsigned long long* tc_trace_buffer(); }
// debug: copy the last BNN_TC_TRACE timeline (3 x 4096 x {tag<<32|it, clock}) to host
extern "C" int32_t bnn_tc_trace_dump(unsigned long long* host_out) {
  unsigned long long* t = bnn::tc_trace_buffer();
  if (!t || !host_out) BNN_FAIL("no trace recorded (needs the trace build, tools/tc_trace.py, and BNN_TC_TRACE=1)");
  BNN_CUDA(cudaMemcpy(host_out, t, 5 * 4096 * 16, cudaMemcpyDeviceToHost));
  return 0;
}

// D[128, N] = A[128, K] * B[N, K]^T on tcgen05 (device pointers; K % 8 == 0, K <= 64; N % 16 == 0, N <= 256)
extern "C" int32_t bnn_tc_selftest(const float* A, const float* B, float* D, int32_t K, int32_t N, int32_t passes, void* stream) {
  if (!A || !B || !D || K < 8 || K > 64 || (K & 7) || N < 16 || N > 256 || (N & 15) || ((passes & 15) < 1 || (passes & 15) > 3) || (passes & ~31) || ((passes & 15) == 2 && (K & 15)))
    BNN_FAIL("bnn_tc_selftest: bad arguments");
  const size_t smem = (size_t)2 * K * 128 * 4 + (size_t)2 * K * N * 4;
  BNN_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 160, smem, (cudaStream_t)stream>>>(A, B, D, K, N, passes);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t bnn_tc_mma_bench(int32_t mode, int32_t N, int32_t iters, long long* dev_out_clk, void* stream) {
  if (mode < 0 || mode > 5 || N < 16 || N > 256 || (N & 15) || iters < 1 || !dev_out_clk) BNN_FAIL("bnn_tc_mma_bench: bad arguments");
  const size_t smem = 48 * 1024;
  BNN_CUDA(cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_mma_bench_kernel<<<1, 160, smem, (cudaStream_t)stream>>>(mode, N, iters, dev_out_clk);
  BNN_CUDA(cudaGetLastError());
  return 0;
}
