// Row f1 of SURVEY.md 8: the whole SGLD minibatch step of BNN_SGDMC.sgld_steps (BNN_SGDMC.py:76-91) as ONE persistent
// cooperative kernel that runs n_steps steps per launch:
//   minibatch gather (DataLoader batch, :80,110) -> L forward layers (util.py:31-33) -> Gaussian log-likelihood + priors and
//   dU/d(out) (:61-74,83) -> backward (dW, dh; what loss.backward() computes, :85) -> preconditioned SGLD update with
//   in-kernel Philox noise and the LambdaLR factor (:86-87,96-101).
// No library GEMM, no elementwise launches, no grid barrier: a step is a list of tile jobs that the CTAs pull from a global
// counter and whose inputs are tracked by completion counters (see "the step as a dataflow job list" below).  The gradient
// never exists in memory: the CTA that finishes a dW tile applies the update to that tile of (theta, V) in its epilogue.
//
// Why FP32 FFMA and not tcgen05 here (north_star: "tensor cores only where width x S makes it a real contraction"): at
// cfg 5 a step is 438 MFLOP spread over a chain of 8 dependent products; one of them is [128 x 512 x 512] = 0.26 MFLOP per
// CTA-tile when spread over 128 SMs, i.e. ~2k cycles of FFMA, the same order as the L2 -> SM ingest of the tile's operands.
// The step is latency / dependency bound, not FLOP bound; FP32 also keeps the gradient bit-comparable (1e-6) with the
// reference's autograd.
//
// Layout: every product is brought into the "NT" form C[m, n] = sum_k A[m, k] * B[n, k] with both operands K-contiguous:
//   forward  l : C = z_l[b, j]        A = h_l[b, 0:ld_l]   (augmented (h | 1 | 0))   B = theta_l[j, 0:ld_l]  ((W | b | 0))
//   dW       l : C = dW_l[j, k]       A = dzT_l[j, 0:Bp]                             B = hT_l[k, 0:Bp]       (row in_l = ones -> db)
//   dh       l : C = dh_l[b, k]       A = dz_l[b, 0:ldz_l]                           B = thetaT_l[k, 0:ldT_l]
// so activations and output gradients are stored twice (row-major and transposed) by the producing epilogue, and the
// weights of layers >= 1 carry a transposed copy maintained by the update epilogue.  theta / thetaT are double-buffered by
// step parity: all reads of a step see theta_cur, all updates write theta_next -- no intra-step hazards.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "philox.cuh"

namespace bnn {

constexpr int kSgThreads = 256;       // compute threads
constexpr int kSgBlock = kSgThreads + 32;   // + the scheduler warp
// compute-only barrier (the scheduler warp never joins): named barrier 1
__device__ __forceinline__ void sg_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
struct SgLayer {
  int in, out, ld, ldz, ldT;
  int ncb, nkb, ndw;         // 16-column tiles of the output / of the input, dW tiles (+ 1 for the last layer: the log-precision job)
  long long off, offT;       // float offsets into theta / thetaT
  float* h[2];               // input of Linear l, augmented: [Bp, ld], by step parity   (l = 0: the gathered batch rows (x | 1 | 0))
  float* hT[2];              // [ld, Bp]
  float* dz;                 // dU/d(pre-activation output of Linear l): [Bp, ldz]
  float* dzT;                // [out, Bp]
  float inv_var;             // 1 / std_l^2 of the weight prior, std_l = gain * sqrt(2 / (in + out))   (BNN_SGDMC.py:65)
};

// one run of same-kind jobs in a step's job list
struct SgSeg { int type, l, first, count; };
constexpr int kSgMaxSeg = 3 * BNN_MAX_LINEAR + 4;
constexpr int kSgMaxRb = 32;           // 32-row blocks of a batch of <= 1024
constexpr int kSgTraceRecs = 96;
constexpr int kSgTraceW = 12;          // words per record       // debug timeline: job records per CTA
// completion counters (unsigned int, zero at launch): [0] = next job index
__host__ __device__ constexpr int sg_c_fwd(int l, int rb) { return 16 + l * kSgMaxRb + rb; }
__host__ __device__ constexpr int sg_c_dh(int l, int rb) { return 16 + (BNN_MAX_LINEAR + l) * kSgMaxRb + rb; }
__host__ __device__ constexpr int sg_c_dht(int l) { return 16 + 2 * BNN_MAX_LINEAR * kSgMaxRb + l; }
__host__ __device__ constexpr int sg_c_dw(int l) { return 16 + 2 * BNN_MAX_LINEAR * kSgMaxRb + BNN_MAX_LINEAR + l; }
__host__ __device__ constexpr int sg_c_ga(int rb) { return 16 + 2 * BNN_MAX_LINEAR * kSgMaxRb + 2 * BNN_MAX_LINEAR + rb; }
constexpr int kSgCounters = 16 + 2 * BNN_MAX_LINEAR * kSgMaxRb + 2 * BNN_MAX_LINEAR + kSgMaxRb;

struct SgParams {
  int L, act, B, Bp, nout, D;
  SgLayer ly[BNN_MAX_LINEAR];
  SgSeg seg[kSgMaxSeg];
  int n_seg, jobs_per_step;
  float* theta[2];
  float* thetaT[2];
  float* V;
  long long n_head, n_theta;           // packed network length (lr_weight group) / total incl. log_precs and padding
  const float* X;
  const float* y;
  const long long* perm;
  long long num_train, perm_off;
  long long t0;                        // global step index of the first step (LR schedule, Philox subsequence)
  int n_steps, cur;
  float lr_w, lr_n, alpha, lam, al, be, lgamma_al, log_be, log_const;
  int precond_mode;                    // 0: G = 1 / (lam + sqrt(V))   1: G = 1 / sqrt(V + lam)
  unsigned long long seed;
  double* acc;                         // by step parity: [nout] sum r^2 per output | prior quad -- zero before launch
  unsigned int* cnt;                   // job / completion counters [kSgCounters], zero before launch
  float* fs;                           // [2] LambdaLR factor of the step, by step parity (written one step ahead by LOGP)
  float f0;                            // ... of the launch's first step
  int* err;
  float* loss_out;                     // U of the LAST step of the launch (may be null)
  float* grad_out;                     // debug: when non-null the step writes dU/dtheta here and does NOT update
  int dbg;                             // debug (BNN_SGLD_DBG): 1 skip the FMA loop, 2 skip the operand loads, 4 skip the epilogues, 8 no noise, 16 no thetaT stores
  long long* trace;                    // debug: [grid][kSgTraceRecs][6] = {job code, grabbed, inputs ready, done (clock64), grabbed, done (globaltimer ns)}, last two steps
};

// ---- small helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float act_fwd(float v, int act) { return apply_act(v, act); }
// derivative of the activation expressed through its OUTPUT h (what autograd's tanh_backward / sigmoid_backward use)
__device__ __forceinline__ float act_bwd(float h, int act) {
  switch (act) {
    case BNN_ACT_RELU: return h > 0.0f ? 1.0f : 0.0f;
    case BNN_ACT_TANH: return 1.0f - h * h;
    case BNN_ACT_SIGMOID: return h * (1.0f - h);
    default: return 1.0f;
  }
}

// ---- operands ----------------------------------------------------------------------------------------------------------
// The kernel is kept SMALL on purpose (rolled loops, ONE inlined copy of the tile core per tile shape, one job loop for all
// phases): a step executes every phase's code exactly once per tile, so straight-line code is instruction-fetch bound -- the
// first version (fully unrolled prefetch prologues, three inlined copies of the core, ~130 KB of SASS) spent 3.5k of a tile's
// 12.9k cycles with the FMAs, the loads and the epilogue all switched off.
struct SgOperand {
  const float* p;      // [rows, ld], K-contiguous
  int ld;
  int rows_valid;      // rows >= rows_valid are not loaded
};

// ---- minibatch gather (DataLoader batch, BNN_SGDMC.py:80,110): 32 rows of the step's batch -> h_0 (augmented (x | 1 | 0), row-
// major) and its transpose.  Done one step AHEAD by a job of the previous step's list (the first step of a launch by a small
// kernel in front), so Linear 0's forward and dW tiles are ordinary dense tiles and the random-row DRAM latency of the gather
// is off the step's critical path.  Rows beyond the live batch (short last batch of an epoch) are zero, ones column included.
__device__ __forceinline__ void sg_gather_rows(const SgParams& p, int it, int rb, float* smem) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const SgLayer& Y = p.ly[0];
  const int par = it & 1, r0 = rb * 32;
  const long long boff = p.perm_off + (long long)it * p.B;
  const long long left = p.num_train - boff;
  const int rows = left < p.B ? (int)left : p.B;
  const int pitch = Y.ld + 1;                        // odd pitch: the transposed read below is conflict-free
  float* h0 = Y.h[par];
  float* h0T = Y.hT[par];
  // warp w takes rows w, w + 8, w + 16, w + 24: the four permutation entries first, then the rows with up to 16 loads in flight
  long long prow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = r0 + w + 8 * i;
    prow[i] = b < rows ? __ldg(p.perm + boff + b) : -1;
  }
#pragma unroll 1
  for (int kk0 = 0; kk0 < Y.ld; kk0 += 128) {
    float v[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = kk0 + u * 32 + lane;
        v[i][u] = 0.0f;
        if (prow[i] >= 0) {
          if (k < p.D) v[i][u] = __ldg(p.X + prow[i] * p.D + k);
          else if (k == p.D) v[i][u] = 1.0f;
        }
      }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = kk0 + u * 32 + lane, r = w + 8 * i;
        if (k < Y.ld) {
          smem[r * pitch + k] = v[i][u];
          if (r0 + r < p.Bp) h0[(long long)(r0 + r) * Y.ld + k] = v[i][u];
        }
      }
  }
  sg_sync();
  if (r0 + lane < p.Bp) {
#pragma unroll 4
    for (int k = w; k < Y.ld; k += kSgThreads / 32) h0T[(long long)k * p.Bp + r0 + lane] = smem[lane * pitch + k];
  }
}

// ---- tile core: red[TM x TN] = sum_k A[m0 + ., k] B[n0 + ., k], left in shared memory (row-major, after the K-slices are summed)
// Operands: every row of the tile arrives WHOLE (its full K extent) through the bulk-copy engine: one cp.async.bulk per row
// and K-half, completion on two mbarriers (transaction bytes).  The first version used 16-byte cp.async per thread: those
// LDGSTS go through the same load/store pipe as the LDS of the FMA loop (194 of them per tile on top of 1,024 LDS.128), and the
// two were additive -- the K loop took 9.8k cycles for 2k cycles of FFMA.  Few large bulk copies matter: the engine retires a
// copy every ~20 cycles whatever its size (432 copies of 256 B per tile took 8k cycles), so rows are not chunked along K.
// Shared-memory row pitch = K rounded up to 32 words + 4: 8 consecutive rows hit 32 distinct banks for float4 reads.
// Rows beyond the operand are not copied (garbage rows only feed output rows / columns the epilogues discard).
// K extents that do not fit shared memory (batch or width > ~800 / ~380) are processed in rounds with a barrier in between.
// Compute: FP32 FFMA, thread tile 8 x 4 (rows ty + TY * i, columns tx + TX * j; 12 LDS.128 per 128 FFMA -- measured 8 % faster per
// step than 4 x 4 with its 16 per 128: the shared-memory pipe is the K loop's bound); the 256 threads form KS = 256 / (TY * TX)
// K-slices, slice ks takes the float4 groups g = ks, ks + KS, ...; slices that share a warp are summed with shuffles.
// Measured alternatives (profiles/r02_sgld_step_notes.md): mma.sync m16n8k8 TF32 with 3xTF32 compensation is NOT faster --
// the legacy warp-level tensor path of this chip issues one TF32 mma per ~32 cycles per SM sub-partition, i.e. the FFMA
// rate, and the compensation needs three of them; 8 x 8 thread tiles (half the LDS per FFMA again) were slower (32 slices to
// reduce, register pressure); 16 warps instead of 8 did not change the K loop; requesting the NEXT job's operands under this
// job's epilogue (look-ahead through the scheduler mailbox) was built and measured slower (64.9 vs 60.5 us per step).
#ifndef BNN_SG_MI
#define BNN_SG_MI 8
#endif
constexpr int kSgMI = BNN_SG_MI;          // rows of a thread's register tile (4 or 8)
constexpr int kSgStageFloats = 39168;     // shared-memory floats for a tile's operands (153 KB)
template <int TM, int TN>
__device__ __forceinline__ void tile_core(const SgOperand& A, const SgOperand& Bop, int m0, int n0, int K,
                                          float* smem, uint64_t* bars, uint32_t& phase_bits, int dbg, long long* probe) {
  constexpr int MI = kSgMI;                          // rows per thread (columns per thread: 4)
  constexpr int TY = TM / MI, TX = TN / 4, THR = TY * TX, KS = kSgThreads / THR, SPW = THR >= 32 ? 1 : 32 / THR;
  constexpr int KRMAX = ((kSgStageFloats / (TM + TN)) - 4) & ~31;
  static_assert(KS >= 1 && KS * THR == kSgThreads && 2 * (TM + TN) <= kSgThreads && (SPW == 1 || SPW * THR == 32), "tile shape");
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = (tid / TX) % TY, ks = tid / THR;
  float acc[MI][4];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  if (probe && tid == 0) probe[0] = clock64();
  const int va = min(TM, A.rows_valid - m0), vb = min(TN, Bop.rows_valid - n0);   // valid rows of the tile (>= 1)
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += KRMAX) {            // one round for every shape of the reference's configurations
    const int Kr = min(KRMAX, K - k0);               // multiple of 4
    const int pitch = ((Kr + 31) & ~31) + 4;
    const int ngr = Kr >> 2;                         // float4 groups along K
    const int nseg = ngr >= 64 ? 2 : 1;              // K-halves with their own barrier: the FMAs start on the first half
    const int gs = nseg == 2 ? ((ngr >> 1) + 3) & ~3 : ngr;
    if (k0 > 0) sg_sync();                           // the previous round's operands are consumed
    if (!(dbg & 2)) {
      if (tid == 0) {
        const uint32_t nrows = (uint32_t)(va + vb);
        mbar_arrive_expect_tx(&bars[0], nrows * (uint32_t)(gs * 16));
        if (nseg == 2) mbar_arrive_expect_tx(&bars[1], nrows * (uint32_t)((ngr - gs) * 16));
      }
      // copy c = row * nseg + half is issued by lane c / 8 of warp c % 8: a warp's lanes issue their bulk copies one after the
      // other (~45 cycles each), so the copies are spread over all eight warps
      {
        const int c = (tid & 31) * (kSgThreads / 32) + (tid >> 5);
        const int row = nseg == 2 ? c >> 1 : c, half = nseg == 2 ? c & 1 : 0;
        if (row < TM + TN) {
          const bool isA = row < TM;
          const int rr = isA ? row : row - TM;
          const SgOperand& op = isA ? A : Bop;
          if (rr < (isA ? va : vb)) {
            const float* src = op.p + (long long)((isA ? m0 : n0) + rr) * op.ld + k0;
            float* dst = smem + row * pitch;
            if (half == 0) bulk_g2s(dst, src, (uint32_t)(gs * 16), &bars[0]);
            else bulk_g2s(dst + gs * 4, src + gs * 4, (uint32_t)((ngr - gs) * 16), &bars[1]);
          }
        }
      }
    }
    if (probe && tid == 0 && k0 == 0) probe[1] = clock64();
    const float* As = smem + ty * pitch;
    const float* Bs = smem + (TM + tx) * pitch;
    auto body = [&](int g) {
      float4 a[MI], b[4];
#pragma unroll
      for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const float4*>(As + TY * i * pitch + g * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bs + TX * j * pitch + g * 4);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    };
    if (!(dbg & 2)) mbar_wait_spin(&bars[0], phase_bits & 1u);
    phase_bits ^= 1u;
    int g = ks;
    if (!(dbg & 1)) {
#pragma unroll 1
      for (; g < gs; g += KS) body(g);
    }
    if (nseg == 2) {
      if (!(dbg & 2)) mbar_wait_spin(&bars[1], (phase_bits >> 1) & 1u);
      phase_bits ^= 2u;
      if (!(dbg & 1)) {
#pragma unroll 1
        for (; g < ngr; g += KS) body(g);
      }
    }
  }
  sg_sync();                               // all slices done with the operands: reuse the memory for the cross-slice reduction
  if (probe && tid == 0) probe[2] = clock64();
  // slices that share a warp are summed with shuffles first; then one partial per remaining slice goes through shared memory
  if (SPW > 1) {
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[i][j];
#pragma unroll
        for (int off = THR; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        acc[i][j] = v;
      }
  }
  constexpr int NP = KS / SPW;             // partials left
  float* red = smem + TM * TN;             // [NP][TM * TN] partials behind the [TM * TN] result
  if (SPW == 1 || (tid & 31) < THR) {
    const int pi = ks / SPW;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[pi * (TM * TN) + (ty + TY * i) * TN + tx + TX * j] = acc[i][j];
  }
  sg_sync();
#pragma unroll 1
  for (int q = tid; q < TM * TN / 4; q += kSgThreads) {
    float4 v = *reinterpret_cast<const float4*>(red + q * 4);
#pragma unroll
    for (int s2 = 1; s2 < NP; ++s2) {
      const float4 w = *reinterpret_cast<const float4*>(red + s2 * (TM * TN) + q * 4);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    *reinterpret_cast<float4*>(smem + q * 4) = v;
  }
  // the next job's bulk copies (async proxy) overwrite what was written here through the generic proxy: order them
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  sg_sync();
  if (probe && tid == 0) probe[3] = clock64();
}

// ---- update of one parameter (shared by the dW epilogue and the log-precision job) ----------------------------------------
__device__ __forceinline__ void psgld_update1(float& t, float& v, float g, float z, float lr, float alpha, float lam, int mode) {
  v = alpha * v + (1.0f - alpha) * g * g;
  const float G = mode == 0 ? 1.0f / (lam + sqrtf(v)) : rsqrtf(v + lam);
  t = t - (0.5f * lr * G * g + sqrtf(lr * G) * z);
}

constexpr int FT_M = 32, FT_N = 16;   // forward / dh tiles over [batch, features]
constexpr int kSgHeadCols = 64;       // columns of the last hidden layer per HEAD job
constexpr int WT_M = 64, WT_N = 32;   // dW tiles over [out, ld]

// ---- the step as a dataflow job list ---------------------------------------------------------------------------------------
// One step = a fixed, topologically ordered list of tile jobs; the launch's list is n_steps copies of it.  CTAs take jobs from
// the head of the list with an atomic counter (the next index is fetched while the current job runs) and wait for the job's
// inputs on monotonically increasing completion counters in global memory (red.release / ld.acquire at gpu scope) -- there is
// no grid-wide barrier anywhere.  Because the list is topologically ordered and every CTA is resident (cooperative launch),
// the oldest unfinished job always has its inputs complete: no deadlock.  Dependencies are per 32-row block of the batch where
// the data flow allows it (forward and dh chains), so the layers of one step overlap, and the weight-update jobs (dW tile +
// prior gradient + pSGLD update) are off the critical path: they only gate the NEXT step's use of that layer.
//   FWD(l, rb, cb)   l <= L-2: h_{l+1}[rb rows, cb cols] = act(h_l theta_l^T)      needs FWD(l-1, rb, *), all DW(l) of step t-1
//   HEAD(rb)         last Linear + Gaussian likelihood head + dz_{L-1} + dz_{L-2} for 32 batch rows (GEMV-sized: no tile core)
//                                                                                needs FWD(L-2, rb, *), all DW(L-1) of step t-1
//   DH(l, rb, kb)    1 <= l <= L-2: dz_{l-1} = (dz_l theta_l) * act'(h_l)          needs DH(l+1, rb, *)  (HEAD(rb) for l = L-2)
//   DW(l, tile)      dW_l tile, + prior gradient, pSGLD update of theta/thetaT/V   needs every producer of dz_l
//   LOGP             log-precision group                                          needs every HEAD
//   LOSS             U of the launch's last step                                  needs everything of that step
// Buffers: theta / thetaT by step parity (readers of step t see "cur", DW writes "next"); h / hT by step parity as well (DW(l+1)
// of step t may still read hT_{l+1} while FWD(l) of step t+1 writes it); dz / dzT and V single (ordered by the chains above).
enum { JT_FWD = 0, JT_HEAD = 1, JT_DH = 2, JT_DW = 3, JT_LOGP = 4, JT_LOSS = 5, JT_GATHER = 6 };

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
__device__ __forceinline__ void red_release_inc(unsigned int* c) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_cta_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// A decoded job, handed from the scheduler warp to the compute warps through shared memory.
struct SgJob {
  int type;                 // JT_*, -1 = the list is exhausted, -2 = a dependency wait timed out
  int l, jj, rb, cb, it;
  int nd, ns;               // counters to wait for / to signal
  int dc[2], sc[2];
  unsigned int dt[2];
  long long t_grab, g_grab, t_ready;   // debug timeline: clock64 / globaltimer at the grab, clock64 when the inputs were complete
};

__device__ __forceinline__ void sg_decode(const SgParams& p, unsigned int idx, unsigned int total, SgJob* J) {
  J->nd = 0; J->ns = 0;
  if (idx >= total) { J->type = -1; return; }
  const int L = p.L;
  const int bm = (p.B + FT_M - 1) / FT_M;
  const int it = (int)(idx / (unsigned int)p.jobs_per_step);
  const int j = (int)(idx - (unsigned int)it * (unsigned int)p.jobs_per_step);
  int si = 0;
  while (si + 1 < p.n_seg && j >= p.seg[si + 1].first) ++si;
  const int type = p.seg[si].type, l = p.seg[si].l, jj = j - p.seg[si].first;
  const SgLayer& Y = p.ly[l];
  const unsigned int done = (unsigned int)it + 1u;                // completions of "this step" = done * per-step count
  int rb = 0, cb = 0, nd = 0, ns = 0;
  switch (type) {
    case JT_GATHER:                                                 // batch rows of step it + 1 (buffer parity of step it - 1)
      rb = jj;
      if (it > 0) { J->dc[nd] = sg_c_dw(0); J->dt[nd++] = (unsigned int)it * (unsigned int)p.ly[0].ndw; }
      J->sc[ns++] = sg_c_ga(rb);
      break;
    case JT_FWD:
      rb = jj / Y.ncb; cb = jj - rb * Y.ncb;
      if (l > 0) { J->dc[nd] = sg_c_fwd(l - 1, rb); J->dt[nd++] = done * (unsigned int)p.ly[l - 1].ncb; }
      else if (it > 0) { J->dc[nd] = sg_c_ga(rb); J->dt[nd++] = (unsigned int)it; }
      if (it > 0) { J->dc[nd] = sg_c_dw(l); J->dt[nd++] = (unsigned int)it * (unsigned int)Y.ndw; }
      J->sc[ns++] = sg_c_fwd(l, rb);
      break;
    case JT_HEAD:
      rb = jj / Y.nkb; cb = jj - rb * Y.nkb;
      if (L > 1) { J->dc[nd] = sg_c_fwd(L - 2, rb); J->dt[nd++] = done * (unsigned int)p.ly[L - 2].ncb; }
      else if (it > 0) { J->dc[nd] = sg_c_ga(rb); J->dt[nd++] = (unsigned int)it; }   // orders the gather before this step's dW_0
      if (it > 0) { J->dc[nd] = sg_c_dw(L - 1); J->dt[nd++] = (unsigned int)it * (unsigned int)Y.ndw; }
      J->sc[ns++] = sg_c_dh(L - 1, rb); J->sc[ns++] = sg_c_dht(L - 1);
      break;
    case JT_DH:
      rb = jj / Y.nkb; cb = jj - rb * Y.nkb;
      J->dc[nd] = sg_c_dh(l + 1, rb); J->dt[nd++] = done * (unsigned int)p.ly[l + 1].nkb;
      J->sc[ns++] = sg_c_dh(l, rb); J->sc[ns++] = sg_c_dht(l);
      break;
    case JT_DW: {
      const int pl = l + 1 < L - 1 ? l + 1 : L - 1;               // who produces dz_l: DH(l + 1), or HEAD for the last two layers
      J->dc[nd] = sg_c_dht(pl); J->dt[nd++] = done * (unsigned int)(bm * p.ly[pl].nkb);
      J->sc[ns++] = sg_c_dw(l);
      break;
    }
    case JT_LOGP:
      J->dc[nd] = sg_c_dht(L - 1); J->dt[nd++] = done * (unsigned int)(bm * Y.nkb);
      J->sc[ns++] = sg_c_dw(L - 1);
      break;
    default: break;                                                 // LOSS waits for everything, see the kernel
  }
  J->type = type; J->l = l; J->jj = jj; J->rb = rb; J->cb = cb; J->it = it; J->nd = nd; J->ns = ns;
}

// Block = 8 compute warps + 1 scheduler warp.  The scheduler warp (one lane) runs one job ahead of the compute warps: it takes
// the next index from the global job counter, decodes it, waits for the job's inputs on the completion counters (ld.acquire at
// gpu scope) and only then posts the job in a two-slot shared-memory mailbox (mbarrier per slot).  The compute warps therefore
// never see the grab / decode / poll latencies (0.6-2.2k cycles per job when done inline): they go from the end of one job to the
// start of the next through one mbarrier wait.  Compute-only synchronisation uses named barrier 1 (256 threads).
__global__ void __launch_bounds__(kSgBlock, 1) sgld_step_kernel(const __grid_constant__ SgParams p) {
  extern __shared__ __align__(16) float sg_smem[];
  __shared__ SgJob s_job[2];
  __shared__ __align__(8) uint64_t s_full[2];              // scheduler -> compute: slot holds a job whose inputs are complete
  __shared__ unsigned int s_go_n, s_done_n;                // compute -> scheduler: jobs started / finished so far
  __shared__ __align__(8) uint64_t s_bars[2];              // "the operands' first / second K-half has landed" (transaction bytes)
  uint32_t phase_bits = 0;             // parity to wait for next, per stage (uniform across the CTA)
  const int tid = threadIdx.x;
  const unsigned int total = (unsigned int)p.n_steps * (unsigned int)p.jobs_per_step;
  const int L = p.L;
  const int bm = (p.B + FT_M - 1) / FT_M;                           // batch-row blocks
  if (tid == 0) {
    mbar_init(&s_bars[0], 1); mbar_init(&s_bars[1], 1);     // one arrival (with the expected bytes) per use
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
    s_go_n = 0; s_done_n = 0;
    fence_barrier_init();
  }
  __syncthreads();

  if (tid >= kSgThreads) {
    // ===================================== scheduler warp =====================================
    if (tid != kSgThreads) return;
    unsigned int lo = 0;                 // oldest handed-over job whose completion has not been published yet
    int bad = 0;
    // publish finished jobs: the compute warps only bump a shared-memory counter at the end of a job; the release at gpu scope
    // (which has to wait for the SM's outstanding stores) is paid here, off their path.  Called from every wait loop.
    auto service = [&]() {
      while (lo < ld_acquire_cta_u32(&s_done_n)) {
        const SgJob* Jd = &s_job[lo & 1u];
        if (Jd->ns > 0) red_release_inc(p.cnt + Jd->sc[0]);
        if (Jd->ns > 1) red_release_inc(p.cnt + Jd->sc[1]);
        ++lo;
      }
    };
    // relaxed polls (an acquire per poll would drop the SM's L1 lines under the compute warps every time), then ONE acquire
    // fence once everything the job needs has been seen complete
    auto wait_for = [&](int c, unsigned int target) {
      unsigned long long spins = 0;
      while ((int)(ld_relaxed_u32(p.cnt + c) - target) < 0) {
        service();
        if (++spins > (1ull << 25)) { atomicExch(p.err, 1); bad = 1; break; }     // never expected: error instead of a hung GPU
        if ((spins & 1023) == 0 && *((volatile int*)p.err) != 0) { bad = 1; break; }
      }
    };
#pragma unroll 1
    for (unsigned int n = 0;; ++n) {
      // one job ahead, not more (load balance): job n is fetched when job n - 1 has started.  Slot n & 1 was last used by job
      // n - 2, which is finished AND published by then (its completion precedes the start of job n - 1).
      if (n > 0) {
        while (ld_acquire_cta_u32(&s_go_n) < n) service();
        service();
      }
      SgJob* J = &s_job[n & 1u];
      const unsigned int idx = atomicAdd(&p.cnt[0], 1u);
      if (p.trace) { J->t_grab = clock64(); J->g_grab = (long long)globaltimer_ns(); }
      sg_decode(p, idx, total, J);
      if (J->type >= 0) {
        for (int d = 0; d < J->nd && !bad; ++d) wait_for(J->dc[d], J->dt[d]);
        if (J->type == JT_LOSS && p.loss_out != nullptr && J->it == p.n_steps - 1) {
          const unsigned int done = (unsigned int)J->it + 1u;
          wait_for(sg_c_dht(L - 1), done * (unsigned int)(bm * p.ly[L - 1].nkb));
          for (int q = 0; q < L && !bad; ++q) wait_for(sg_c_dw(q), done * (unsigned int)p.ly[q].ndw);
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      if (bad) J->type = -2;
      if (p.trace) J->t_ready = clock64();
      const int ty = J->type;
      mbar_arrive(&s_full[n & 1u]);                                 // releases the slot's contents to the compute warps
      if (ty < 0) {
        // drain: publish what is still running (all of it finishes: compute warps never wait for this warp except for new jobs)
        if (ty == -1) while (lo < n) service();
        return;
      }
    }
  }

  // ======================================== compute warps ========================================
  const bool debug = p.grad_out != nullptr;
  int trec = 0;
  long long t_prev_end = 0;
#pragma unroll 1
  for (unsigned int n = 0;; ++n) {
    long long t_start = 0, pr[4] = {0, 0, 0, 0};
    const SgJob* J = &s_job[n & 1u];
    mbar_wait_spin(&s_full[n & 1u], (n >> 1) & 1u);
    const int type = J->type;
    if (type < 0) return;
    const int l = J->l, jj = J->jj, rb = J->rb, cb = J->cb, it = J->it;
    const long long t_grab = J->t_grab, g_grab = J->g_grab, t_ready = J->t_ready;
    sg_sync();                                                      // everybody has read the slot
    if (tid == 0) st_release_cta_u32(&s_go_n, n + 1u);
    if (p.trace && tid == 0) t_start = clock64();

    const long long t = p.t0 + it;
    const int cur = (p.cur + it) & 1, par = it & 1;
    const float* th = p.theta[cur];
    const float* thT = p.thetaT[cur];
    float* th_next = p.theta[cur ^ 1];
    float* thT_next = p.thetaT[cur ^ 1];
    const long long boff = p.perm_off + (long long)it * p.B;
    const long long left = p.num_train - boff;
    const int rows = left < p.B ? (int)left : p.B;                 // the last batch of an epoch may be short (:80)
    const float scale = (float)((double)p.num_train / (double)rows);   // N / |b|  (:83)
    const bool want_loss = p.loss_out != nullptr && it == p.n_steps - 1;
    double* acc = p.acc + par * (p.nout + 1);
    const SgLayer& Y = p.ly[l];

    // =================================== part A: the job's main work ===================================
    const bool tile_job = type == JT_DW || type == JT_FWD || type == JT_DH;
    const bool big = type == JT_DW && l > 0;                        // 64 x 32 dW tiles; Linear 0's dW (gathered operand, on the
    const int tm = big ? WT_M : FT_M, tn_sh = big ? 5 : 4;          // critical path into the next step) uses 32 x 16 tiles
    int m0 = 0, n0 = 0;
    if (tile_job) {
      SgOperand A, Bo;
      int K;
      if (type == JT_DW) {
        // dW_l tile over K = batch
        const int wm = (Y.out + tm - 1) / tm;
        m0 = (jj % wm) * tm; n0 = (jj / wm) << tn_sh;
        A.p = Y.dzT; A.ld = p.Bp; A.rows_valid = Y.out;
        Bo.p = Y.hT[par]; Bo.ld = p.Bp; Bo.rows_valid = Y.ld;
        K = p.Bp;
      } else if (type == JT_FWD) {
        m0 = rb * FT_M; n0 = cb * FT_N;
        A.p = Y.h[par]; A.ld = Y.ld; A.rows_valid = p.B;
        Bo.p = th + Y.off; Bo.ld = Y.ld; Bo.rows_valid = Y.out;
        K = Y.ld;
      } else {
        m0 = rb * FT_M; n0 = cb * FT_N;
        A.p = Y.dz; A.ld = Y.ldz; A.rows_valid = p.B;
        Bo.p = thT + Y.offT; Bo.ld = Y.ldT; Bo.rows_valid = Y.in;
        K = Y.ldz;
      }
      if (big) tile_core<WT_M, WT_N>(A, Bo, m0, n0, K, sg_smem, s_bars, phase_bits, p.dbg, p.trace ? pr : nullptr);
      else tile_core<FT_M, FT_N>(A, Bo, m0, n0, K, sg_smem, s_bars, phase_bits, p.dbg, p.trace ? pr : nullptr);
    }
    // HEAD(rb, chunk): last Linear, residual, dU/d(out), per-output sum of squares (BNN_SGDMC.py:69-74, :83) for 32 batch rows,
    // then dz of the last hidden layer for a 64-column chunk of those rows.  GEMV-sized work (nout is 1..few): eight threads per
    // row along K, no tile core.  The (cheap) first part is recomputed by every chunk's job so that the second part spreads over
    // bm x (width / 64) CTAs without another dependency hop; chunk 0 owns the side effects (dz_{L-1}, sums of squares).
    const SgLayer& Yl = p.ly[L - 1];
    float* s_dz = sg_smem;                                          // [FT_M][nout]
    float* s_w = sg_smem + ((FT_M * p.nout + 3) & ~3);              // [nout][ld]: the last Linear's rows
    float* s_h = s_w + p.nout * Yl.ld;                              // [FT_M][ld]: this row block of the last hidden layer
    if (type == JT_HEAD) {
      const int r0 = rb * FT_M;
      const int nw = p.nout * Y.ld;
      // everything the job reads from other SMs' output arrives in ONE round trip: 16-byte cp.async for W and the h rows
      for (int i = tid * 4; i < nw; i += kSgThreads * 4) cp_async16(smem_u32(s_w + i), th + Y.off + i, 16);
      if (L > 1) {
        const int q4 = Y.ld >> 2;
        const float* hsrc = Y.h[par] + (long long)r0 * Y.ld;
        const int nrow = min(FT_M, p.B - r0);
        for (int i = tid; i < nrow * q4; i += kSgThreads) cp_async16(smem_u32(s_h + i * 4), hsrc + (long long)i * 4, 16);
      }
      cp_async_commit();
      const int row = tid >> 3, sub = tid & 7, b = r0 + row;
      const bool live = b < rows;
      long long prow = 0;
      float ytrue = 0.0f;
      if (live) {
        prow = __ldg(p.perm + boff + b);
        if (sub < p.nout) ytrue = __ldg(p.y + prow * p.nout + sub);   // targets of outputs 0..7 travel with the operands
      }
      cp_async_wait<0>();
      sg_sync();
      if (p.trace && tid == 0) pr[0] = clock64();
      const float* hin = s_h + row * Y.ld;
#pragma unroll 1
      for (int o = 0; o < p.nout; ++o) {
        const float* w = s_w + o * Y.ld;
        float a = 0.0f;
        if (b < p.B) {
          if (L > 1) {
#pragma unroll 4
            for (int k = sub * 4; k < Y.ld; k += 32) {
              const float4 hv = *reinterpret_cast<const float4*>(hin + k);
              const float4 wv = *reinterpret_cast<const float4*>(w + k);
              a = fmaf(hv.x, wv.x, a); a = fmaf(hv.y, wv.y, a); a = fmaf(hv.z, wv.z, a); a = fmaf(hv.w, wv.w, a);
            }
          } else if (live) {
            for (int k = sub; k <= p.D; k += 8) a = fmaf(k < p.D ? __ldg(p.X + prow * p.D + k) : 1.0f, w[k], a);
          }
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        // the row's target for output o sits in lane (o & 7) of the row's eight lanes for o < 8
        float yt = __shfl_sync(0xffffffffu, ytrue, (tid & 24) | (o & 7));
        if (o >= 8 && live) yt = __ldg(p.y + prow * p.nout + o);
        float dzv = 0.0f;
        double r2 = 0.0;
        if (sub == 0) {
          if (live) {
            const float r = a - yt;
            dzv = scale * expf(th[p.n_head + o]) * r;
            r2 = (double)r * (double)r;
          }
          if (b < p.B && cb == 0) {
            Y.dz[(long long)b * Y.ldz + o] = dzv;
            Y.dzT[(long long)o * p.Bp + b] = dzv;
          }
          s_dz[row * p.nout + o] = dzv;
        }
#pragma unroll
        for (int q = 16; q > 0; q >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, q);
        if ((tid & 31) == 0 && r2 != 0.0 && cb == 0) atomicAdd(&acc[o], r2);
      }
      if (p.trace && tid == 0) pr[1] = clock64();
      sg_sync();
    }

    // =================================== part B: epilogues ===================================
    if (type == JT_DW) {
      // ---- prior gradient + pSGLD update of the tile of theta / V (the gradient itself is never stored)
      double quad_local = 0.0;
      if (!(p.dbg & 4)) {
        const float lr_w = p.lr_w * (it == 0 ? p.f0 : p.fs[par]);    // LambdaLR factor of this step (:101)
        const int nquads = (tm << tn_sh) >> 2, tn_mask = (1 << tn_sh) - 1;
        // a thread owns up to two quads of the tile (q = tid, tid + 256): both quads' theta / V loads are issued before any
        // arithmetic so their L2 latency is paid once
        float4 t4[2], v4[2];
        bool on[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int q = tid + u * kSgThreads;
          const int jr = m0 + ((q * 4) >> tn_sh), k0 = n0 + ((q * 4) & tn_mask);
          on[u] = q < nquads && jr < Y.out && k0 < Y.ld;
          if (on[u]) {
            const long long e = Y.off + (long long)jr * Y.ld + k0;
            // plain (L1-allocating) loads are coherent here: every job starts behind an acquire at gpu scope, which drops
            // this SM's stale L1 lines; ld.global.cg (LDG.STRONG.GPU) costs ~3x the latency on this part
            t4[u] = *reinterpret_cast<const float4*>(th + e);
            if (!debug) v4[u] = *reinterpret_cast<const float4*>(p.V + e);
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (!on[u]) continue;
          const int q = tid + u * kSgThreads;
          const int jr = m0 + ((q * 4) >> tn_sh), k0 = n0 + ((q * 4) & tn_mask);
          const float4 v = *reinterpret_cast<const float4*>(sg_smem + q * 4);
          const long long e = Y.off + (long long)jr * Y.ld + k0;         // multiple of 4: one Philox block per quad
          float tt[4] = {t4[u].x, t4[u].y, t4[u].z, t4[u].w};
          float g[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int k = k0 + c;
            if (k < Y.in) {                                                // weight entry: + d(-log prior)/dW = W / std^2  (:65-66)
              g[c] = fmaf(Y.inv_var, tt[c], g[c]);
              quad_local += (double)(Y.inv_var * tt[c]) * (double)tt[c];
            } else if (k > Y.in) g[c] = 0.0f;                              // layout padding: no gradient, no noise, stays zero
          }
          if (debug) {
            *reinterpret_cast<float4*>(p.grad_out + e) = make_float4(g[0], g[1], g[2], g[3]);
            continue;
          }
          float vv[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
          float z[4];
          if (p.dbg & 8) { z[0] = z[1] = z[2] = z[3] = 0.0f; }
          else philox_normal4(p.seed, (uint32_t)t, STREAM_SGLD, (uint64_t)(e >> 2), z);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int k = k0 + c;
            if (k <= Y.in) psgld_update1(tt[c], vv[c], g[c], z[c], lr_w, p.alpha, p.lam, p.precond_mode);
            else tt[c] = 0.0f;
          }
          *reinterpret_cast<float4*>(th_next + e) = make_float4(tt[0], tt[1], tt[2], tt[3]);
          *reinterpret_cast<float4*>(p.V + e) = make_float4(vv[0], vv[1], vv[2], vv[3]);
          if (l > 0 && !(p.dbg & 16)) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int k = k0 + c;
              if (k < Y.in) thT_next[Y.offT + (long long)k * Y.ldT + jr] = tt[c];
            }
          }
        }
      }
      if (want_loss) {
        // prior energy sum W^2 / std^2 of the step's INPUT parameters (the loss is evaluated before the update, :81-83)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) quad_local += __shfl_xor_sync(0xffffffffu, quad_local, o);
        if ((tid & 31) == 0 && quad_local != 0.0) atomicAdd(&acc[p.nout], quad_local);
      }
    } else if (type == JT_FWD || type == JT_DH) {
      if (!(p.dbg & 4) && tid < FT_M * FT_N / 4) {
        const int b = m0 + (tid * 4) / FT_N, c0 = n0 + (tid * 4) % FT_N;
        const float4 v = *reinterpret_cast<const float4*>(sg_smem + tid * 4);
        const float vv[4] = {v.x, v.y, v.z, v.w};
        if (b < p.B) {
          if (type == JT_FWD) {
            // hidden layer: h_{l+1} = act(z), stored row-major and transposed
            const SgLayer& Nx = p.ly[l + 1];
            float* hn = Nx.h[par];
            float* hnT = Nx.hT[par];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int jc = c0 + c;
              if (jc < Y.out) {
                const float a = act_fwd(vv[c], p.act);
                hn[(long long)b * Nx.ld + jc] = a;
                hnT[(long long)jc * p.Bp + b] = a;
              }
            }
          } else {
            // dz_{l-1} = dh * act'(h_l)
            const SgLayer& Pv = p.ly[l - 1];
            const float* hl = Y.h[par];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int k = c0 + c;
              if (k < Y.in) {
                const float hv = hl[(long long)b * Y.ld + k];
                const float d = vv[c] * act_bwd(hv, p.act);
                Pv.dz[(long long)b * Pv.ldz + k] = d;
                Pv.dzT[(long long)k * p.Bp + b] = d;
              }
            }
          }
        }
      }
    } else if (type == JT_HEAD) {
      if (L > 1) {
        // dz_{L-2}[b, k] = (sum_o dz_{L-1}[b, o] W[o, k]) * act'(h[b, k]) for k in this job's 64-column chunk.  Pass 1: thread =
        // (column, 8-row group), coalesced row-major stores, the value also replaces h in shared memory; pass 2: lane = batch
        // row, so the transposed copy goes out as full 128-byte rows too.
        const int r0 = rb * FT_M, kc0 = cb * kSgHeadCols;
        const SgLayer& Pv = p.ly[L - 2];
        const int nrow = min(FT_M, p.B - r0);
        {
          const int k = kc0 + (tid & (kSgHeadCols - 1)), u0 = (tid / kSgHeadCols) * 8;
          if (k < Y.in) {
            float* dzp = Pv.dz + (long long)r0 * Pv.ldz + k;
            float wk[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) wk[o] = o < p.nout ? s_w[o * Y.ld + k] : 0.0f;
#pragma unroll
            for (int uu = 0; uu < 8; ++uu) {
              const int u = u0 + uu;
              float dh = s_dz[u * p.nout] * wk[0];
              if (p.nout > 1) {
#pragma unroll
                for (int o = 1; o < 4; ++o) if (o < p.nout) dh = fmaf(s_dz[u * p.nout + o], wk[o], dh);
#pragma unroll 1
                for (int o = 4; o < p.nout; ++o) dh = fmaf(s_dz[u * p.nout + o], s_w[o * Y.ld + k], dh);
              }
              const float d = u < nrow ? dh * act_bwd(s_h[u * Y.ld + k], p.act) : 0.0f;
              s_h[u * Y.ld + k] = d;
              if (u < nrow) dzp[(long long)u * Pv.ldz] = d;
            }
          }
        }
        sg_sync();
        const int lane = tid & 31;
        if (r0 + lane < p.Bp) {
#pragma unroll
          for (int kk = tid >> 5; kk < kSgHeadCols; kk += kSgThreads / 32) {
            const int k = kc0 + kk;
            if (k < Y.in) Pv.dzT[(long long)k * p.Bp + r0 + lane] = s_h[lane * Y.ld + k];
          }
        }
      }
    } else if (type == JT_GATHER) {
      if (it + 1 < p.n_steps) sg_gather_rows(p, it + 1, rb, sg_smem);
    } else if (type == JT_LOGP) {
      // ---- log-precision group (BNN_SGDMC.py:62, :96-98): closed-form gradient, same update rule, lr_noise
      if (tid < ((p.n_theta - p.n_head) >> 2)) {
        const float lr_n = p.lr_n * (it == 0 ? p.f0 : p.fs[par]);
        const long long e = p.n_head + (long long)tid * 4;
        float tt[4], vv[4], g[4], z[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tt[c] = th[e + c];
          vv[c] = p.V[e + c];
          g[c] = 0.0f;
          const long long o = e + c - p.n_head;
          if (o < p.nout) {
            const float prec = expf(tt[c]);
            const float s2 = (float)acc[o];
            g[c] = -scale * (-0.5f * prec * s2 + 0.5f * (float)rows) - (p.al - p.be * prec);
          }
        }
        if (debug) {
#pragma unroll
          for (int c = 0; c < 4; ++c) p.grad_out[e + c] = g[c];
        } else {
          philox_normal4(p.seed, (uint32_t)t, STREAM_SGLD, (uint64_t)(e >> 2), z);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (e + c - p.n_head < p.nout) psgld_update1(tt[c], vv[c], g[c], z[c], lr_n, p.alpha, p.lam, p.precond_mode);
            else tt[c] = 0.0f;
            th_next[e + c] = tt[c];
            p.V[e + c] = vv[c];
          }
        }
      }
      // the other parity's sums of squares were consumed by the previous step's LOGP (which every HEAD of this step waited
      // for) and are next written by the HEADs of step t + 1, which wait for this job: reset them here
      if (tid < p.nout) p.acc[(par ^ 1) * (p.nout + 1) + tid] = 0.0;
      // LambdaLR: np.float32((1 + it) ** -0.33) of the NEXT step, in double on one thread; its readers (the update jobs of
      // step t + 1) are behind this job, the slot's previous readers (step t - 1) before it
      if (tid == 64) p.fs[par ^ 1] = (float)pow((double)(2 + t), -0.33);
    } else if (type == JT_LOSS) {
      if (want_loss && tid == 0) {
        float total_ll = 0.0f, total_lp = 0.0f;
        for (int o = 0; o < p.nout; ++o) {
          const float lp = th[p.n_head + o], prec = expf(lp);
          const float s2 = (float)acc[o];
          total_ll += -0.5f * prec * s2 + (float)rows * (0.5f * lp - 0.91893853320467274178f);
          total_lp += p.al * p.log_be + (p.al - 1.0f) * lp - p.be * prec - p.lgamma_al + lp;
        }
        const float quad = (float)acc[p.nout];
        const float logp = -0.5f * quad - p.log_const + total_lp;
        p.loss_out[0] = -(total_ll * scale + logp);
      }
    }

    // ---- done: tell the scheduler warp, which publishes the completion.  The next job's bulk copies (async proxy) overwrite
    // shared memory this job wrote through the generic proxy: order them.  (Global memory needs no proxy fence here: the bulk
    // copies read L2, where the producers' stores are before their completion is published with release / acquire at gpu scope;
    // a full fence.proxy.async on either side was measured at 2.5-5k cycles per job.)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    sg_sync();                           // the job's global writes are issued; its shared memory is free for the next job
    if (tid == 0) {
      st_release_cta_u32(&s_done_n, n + 1u);
      if (p.trace) {
        const long long t_end = clock64();
        if (it >= p.n_steps - 2) {
          long long* r = p.trace + ((long long)blockIdx.x * kSgTraceRecs + (trec < kSgTraceRecs ? trec : kSgTraceRecs - 1)) * kSgTraceW;
          r[0] = ((long long)it << 32) | ((long long)type << 24) | ((long long)l << 16) | (long long)(jj & 0xffff);
          r[1] = t_grab; r[2] = t_start; r[3] = t_end; r[4] = g_grab; r[5] = (long long)globaltimer_ns();
          r[6] = pr[0]; r[7] = pr[1]; r[8] = t_ready; r[9] = t_prev_end; r[10] = pr[2]; r[11] = pr[3];
          ++trec;
        }
        t_prev_end = t_end;
      }
    }
  }
}

// batch rows of a launch's FIRST step (later steps are gathered by jobs of the step before)
__global__ void __launch_bounds__(kSgThreads, 1) sgld_gather_kernel(const __grid_constant__ SgParams p) {
  extern __shared__ __align__(16) float sg_smem[];
  sg_gather_rows(p, 0, blockIdx.x, sg_smem);
}

// theta[cur] -> thetaT[cur] (layers >= 1), once per chain initialisation
__global__ void sgld_transpose_kernel(const __grid_constant__ SgParams p, int buf) {
  for (int l = 1; l < p.L; ++l) {
    const SgLayer& Y = p.ly[l];
    const long long n = (long long)Y.out * Y.in;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const int j = (int)(i / Y.in), k = (int)(i - (long long)j * Y.in);
      p.thetaT[buf][Y.offT + (long long)k * Y.ldT + j] = p.theta[buf][Y.off + (long long)j * Y.ld + k];
    }
  }
}

// ones column / ones row of the augmented activations
__global__ void sgld_ones_kernel(const __grid_constant__ SgParams p) {
  for (int l = 1; l < p.L; ++l) {
    const SgLayer& Y = p.ly[l];
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < p.B; b += gridDim.x * blockDim.x) {
      for (int q = 0; q < 2; ++q) {
        Y.h[q][(long long)b * Y.ld + Y.in] = 1.0f;
        Y.hT[q][(long long)Y.in * p.Bp + b] = 1.0f;
      }
    }
  }
}

// ---- epoch permutation (DataLoader(shuffle=True), BNN_SGDMC.py:110) -------------------------------------------------------
// A chain consumes only the first n_steps * batch entries of an epoch's permutation (cfg 5: 6,400 of 10 M per thinning
// interval), so the permutation is never materialised: entry i is computed directly as pi(i), with pi a keyed bijection of
// [0, N) -- a balanced Feistel network over 2^(2 hb) >= N (round function: one Philox4x32-10 block keyed by (seed, epoch))
// with cycle walking (re-encrypt until the value falls inside [0, N); the domain is < 4 N, so < 4 rounds on average).
__device__ __forceinline__ unsigned long long feistel_perm(unsigned long long x, int hb, unsigned long long seed, uint32_t epoch) {
  const uint32_t mask = hb >= 32 ? 0xffffffffu : ((1u << hb) - 1u);
  uint32_t lo = (uint32_t)x & mask, hi = (uint32_t)(x >> hb) & mask;
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
    const u32x4 f = philox4x32_10(lo, (uint32_t)r, epoch, 0x5e1fu, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t nh = lo, nl = (hi ^ f.x) & mask;
    hi = nh; lo = nl;
  }
  return ((unsigned long long)hi << hb) | lo;
}
__global__ void sgld_perm_kernel(unsigned long long seed, uint32_t epoch, long long N, long long off, long long count, int hb,
                                 long long* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  unsigned long long x = (unsigned long long)(off + i);
  do { x = feistel_perm(x, hb, seed, epoch); } while (x >= (unsigned long long)N);
  out[i] = (long long)x;
}

// ---- host side -------------------------------------------------------------------------------------------------------------
static inline long long r4(long long v) { return (v + 3) & ~3LL; }

struct SgPlan {
  SgParams p;
  long long nT;            // floats of ONE thetaT buffer
  long long ws_floats;     // activation workspace (floats), followed by acc / barrier / err
  long long ws_bytes, tail_bytes;
  int grid;
  size_t smem;
};

static int sg_plan(const BnnSgldChain* c, SgPlan* plan, bool bind, bool want_loss = false) {
  Layout Lo;
  if (make_layout(&c->desc, &Lo)) return -1;
  if (c->batch < 1 || c->batch > 1024) BNN_FAIL("bnn_sgld: batch must be in [1, 1024] (got %d)", c->batch);
  SgParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.L = Lo.n_lin; p.act = Lo.act; p.B = c->batch; p.Bp = (int)r4(c->batch); p.nout = Lo.out[Lo.n_lin - 1]; p.D = Lo.in[0];
  if (p.nout > BNN_MAX_FUSED_NOUT) BNN_FAIL("bnn_sgld: nout=%d exceeds %d", p.nout, BNN_MAX_FUSED_NOUT);
  p.n_head = Lo.off[Lo.n_lin];
  p.n_theta = r4(p.n_head + p.nout);
  long long oT = 0, ow = 0;
  double log_const = 0.0;
  const double gain = c->gain > 0.f ? c->gain : 5.0 / 3.0;
  const int bm = (p.B + FT_M - 1) / FT_M;
  for (int l = 0; l < p.L; ++l) {
    SgLayer& Y = p.ly[l];
    Y.in = Lo.in[l]; Y.out = Lo.out[l]; Y.ld = Lo.ld[l]; Y.off = Lo.off[l];
    Y.ldz = (int)r4(Y.out); Y.ldT = (int)r4(Y.out);
    Y.offT = oT;
    if (l > 0) oT += (long long)Y.in * Y.ldT;
    const double std = gain * sqrt(2.0 / (double)(Y.in + Y.out));
    Y.inv_var = (float)(1.0 / (std * std));
    log_const += (double)Y.in * Y.out * (log(std) + 0.5 * log(2.0 * 3.14159265358979323846));
    Y.ncb = (Y.out + FT_N - 1) / FT_N;
    Y.nkb = (Y.in + FT_N - 1) / FT_N;
    if (l == p.L - 1) Y.nkb = l > 0 ? (Y.in + kSgHeadCols - 1) / kSgHeadCols : 1;   // last layer: column chunks of the HEAD jobs
    // dW tiles: 64 x 32, except Linear 0 (32 x 16: its update is on the critical path into the next step)
    Y.ndw = (l > 0 ? ((Y.out + WT_M - 1) / WT_M) * ((Y.ld + WT_N - 1) / WT_N) : ((Y.out + FT_M - 1) / FT_M) * ((Y.ld + FT_N - 1) / FT_N)) +
            (l == p.L - 1 ? 1 : 0);
  }
  plan->smem = (size_t)kSgStageFloats * 4;
  // the head job keeps the last Linear's rows, 32 rows of the last hidden layer and of dU/d(out) in shared memory
  if ((size_t)32 * (p.ly[0].ld + 1) > (size_t)kSgStageFloats)
    BNN_FAIL("bnn_sgld: input dimension %d is too large for the in-kernel batch gather", p.D);
  {
    const size_t head = ((size_t)FT_M * p.nout + 4 + (size_t)(p.nout + FT_M) * p.ly[p.L - 1].ld) * 4;
    if (head > (size_t)227 * 1024)
      BNN_FAIL("bnn_sgld: (nout + 32) * (last hidden width) = %d * %d does not fit the fused head's shared memory", p.nout + FT_M, p.ly[p.L - 1].ld);
    if (head > plan->smem) plan->smem = head;
  }
  // the step's job list (topological; see the kernel): forward chain, head, then dh chain with the updates slotted behind
  // the dh segment that runs while their input is being finished
  {
    int n = 0, first = 0;
    auto add = [&](int type, int l, int count) {
      if (count <= 0) return;
      p.seg[n].type = type; p.seg[n].l = l; p.seg[n].first = first; p.seg[n].count = count;
      first += count; ++n;
    };
    const int Ln = p.L;
    auto n_dw = [&](int l) { return p.ly[l].ndw - (l == Ln - 1 ? 1 : 0); };
    for (int l = 0; l + 1 < Ln; ++l) add(JT_FWD, l, bm * p.ly[l].ncb);
    add(JT_HEAD, Ln - 1, bm * p.ly[Ln - 1].nkb);
    if (Ln <= 2) add(JT_GATHER, 0, bm);
    for (int l = Ln - 2; l >= 1; --l) {
      add(JT_DH, l, bm * p.ly[l].nkb);
      if (l == Ln - 2) add(JT_GATHER, 0, bm);            // the next step's batch rows, off the critical path
      add(JT_DW, l + 1, n_dw(l + 1));
      if (l + 1 == Ln - 1) add(JT_LOGP, Ln - 1, 1);
    }
    add(JT_DW, 0, n_dw(0));
    if (Ln >= 2) add(JT_DW, 1, n_dw(1));
    if (Ln <= 2) add(JT_LOGP, Ln - 1, 1);
    if (want_loss) add(JT_LOSS, 0, 1);
    p.n_seg = n; p.jobs_per_step = first;
  }
  plan->nT = oT > 0 ? oT : 4;
  // workspace: per layer h, hT (two copies, by step parity; layer 0: the gathered batch), dz, dzT
  for (int l = 0; l < p.L; ++l) {
    SgLayer& Y = p.ly[l];
    ow += 2 * (r4((long long)p.Bp * Y.ld) + r4((long long)Y.ld * p.Bp));
    ow += r4((long long)p.Bp * Y.ldz);
    ow += r4((long long)Y.out * p.Bp);
  }
  plan->ws_floats = ow;
  plan->tail_bytes = 16LL * (p.nout + 1) + 4LL * kSgCounters + 64;   // acc doubles (two parities), counters, fs[2], error flag
  plan->ws_bytes = ((ow * 4 + 15) & ~15LL) + plan->tail_bytes;
  p.log_const = (float)log_const;
  p.lr_w = c->lr_weight; p.lr_n = c->lr_noise; p.alpha = c->precond_decay; p.lam = c->diagonal_bias;
  p.al = c->alpha_n; p.be = c->beta_n; p.lgamma_al = lgammaf(c->alpha_n); p.log_be = logf(c->beta_n);
  p.precond_mode = c->precond_mode;
  p.seed = c->seed;
  int grid = sm_count();
  if (p.jobs_per_step < grid) grid = p.jobs_per_step;
  plan->grid = grid;
  if (bind) {
    if (!c->theta || !c->thetaT || !c->V || !c->workspace) BNN_FAIL("bnn_sgld: null chain buffers");
    if (c->workspace_bytes < plan->ws_bytes) BNN_FAIL("bnn_sgld: workspace needs %lld bytes, got %lld", plan->ws_bytes, (long long)c->workspace_bytes);
    if ((((uintptr_t)c->theta | (uintptr_t)c->thetaT | (uintptr_t)c->V | (uintptr_t)c->workspace) & 15) != 0) BNN_FAIL("bnn_sgld: buffers must be 16-byte aligned");
    p.theta[0] = c->theta; p.theta[1] = c->theta + p.n_theta;
    p.thetaT[0] = c->thetaT; p.thetaT[1] = c->thetaT + plan->nT;
    p.V = c->V;
    float* w = reinterpret_cast<float*>(c->workspace);
    long long o = 0;
    for (int l = 0; l < p.L; ++l) {
      SgLayer& Y = p.ly[l];
      for (int q = 0; q < 2; ++q) { Y.h[q] = w + o; o += r4((long long)p.Bp * Y.ld); Y.hT[q] = w + o; o += r4((long long)Y.ld * p.Bp); }
      Y.dz = w + o; o += r4((long long)p.Bp * Y.ldz);
      Y.dzT = w + o; o += r4((long long)Y.out * p.Bp);
    }
    char* tailp = reinterpret_cast<char*>(c->workspace) + ((ow * 4 + 15) & ~15LL);
    p.acc = reinterpret_cast<double*>(tailp);
    p.cnt = reinterpret_cast<unsigned int*>(tailp + 16LL * (p.nout + 1));
    p.fs = reinterpret_cast<float*>(tailp + 16LL * (p.nout + 1) + 4LL * kSgCounters);
    p.err = reinterpret_cast<int*>(tailp + 16LL * (p.nout + 1) + 4LL * kSgCounters + 16);
  }
  return 0;
}

}  // namespace bnn

using namespace bnn;

extern "C" int64_t bnn_sgld_theta_len(const BnnMlpDesc* d) {
  Layout Lo;
  if (make_layout(d, &Lo)) return -1;
  return r4(Lo.off[Lo.n_lin] + Lo.out[Lo.n_lin - 1]);
}

extern "C" int64_t bnn_sgld_theta_t_len(const BnnMlpDesc* d) {
  BnnSgldChain c;
  memset(&c, 0, sizeof(c));
  if (!d) { set_error("null desc"); return -1; }
  c.desc = *d; c.batch = 4; c.alpha_n = 1.f; c.beta_n = 1.f;
  SgPlan plan;
  if (sg_plan(&c, &plan, false)) return -1;
  return plan.nT;
}

extern "C" int64_t bnn_sgld_workspace_bytes(const BnnMlpDesc* d, int32_t batch) {
  BnnSgldChain c;
  memset(&c, 0, sizeof(c));
  if (!d) { set_error("null desc"); return -1; }
  c.desc = *d; c.batch = batch; c.alpha_n = 1.f; c.beta_n = 1.f;
  SgPlan plan;
  if (sg_plan(&c, &plan, false)) return -1;
  return plan.ws_bytes;
}

// zero the workspace, set the ones column / row of the augmented activations, build thetaT[cur] from theta[cur]
extern "C" int32_t bnn_sgld_init(const BnnSgldChain* c, int32_t cur, void* stream) {
  if (!c || (cur != 0 && cur != 1)) BNN_FAIL("bnn_sgld_init: bad arguments");
  SgPlan plan;
  if (sg_plan(c, &plan, true)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  BNN_CUDA(cudaMemsetAsync(c->workspace, 0, (size_t)plan.ws_bytes, st));
  BNN_CUDA(cudaMemsetAsync(c->thetaT, 0, (size_t)plan.nT * 2 * 4, st));
  sgld_ones_kernel<<<4, 256, 0, st>>>(plan.p);
  sgld_transpose_kernel<<<64, 256, 0, st>>>(plan.p, cur);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- the same step for networks that fit ONE SM ----------------------------------------------------------------------------
// The reference's own SGLD use (experiments/SGDMC.py, BO.py: one hidden layer of 50 units, batch 32) is a 751-parameter chain:
// a step is ~150 kFLOP, and on the dataflow kernel above it costs seven inter-SM hand-overs (~15 us).  When theta, V, the
// batch's activations and two gradient panels fit one SM's shared memory, `sgld_small_kernel` runs the launch's n steps as ONE
// CTA with the chain state resident in shared memory: gather -> forward -> Gaussian head -> per layer (input gradient, then
// dW + prior gradient + pSGLD update of the thread's own entries) -> log-precision group, CTA barriers only.  Same arithmetic,
// same Philox stream (seed, t, block = packed offset / 4) and same LambdaLR factors as the dataflow kernel, so the two are
// interchangeable between launches (BNN_SGLD_SMALL=0 forces the dataflow kernel; tests compare them step for step).
// Thread tiles as in csrc/dropout_train.cu: weight / activation rows at a pitch lp with lp / 4 odd, gradient panels transposed.
constexpr int kSsThreads = 512;          // 14 compute warps + 2 prefetch warps
constexpr int kSsPrefetch = 64;
constexpr int kSsCompute = kSsThreads - kSsPrefetch;
constexpr long long kSsMaxWork = 96 * 1024;
struct SsLayer { int lp, half, k4p, soff, a_off; };
struct SsParams {
  SsLayer sy[BNN_MAX_LINEAR];
  int sP, gp, max_w, hd;                 // hd = n_theta - n_head (log-precisions + pad)
  int sm_v, sm_z, sm_a0, sm_act, sm_g0, sm_g1, sm_y, sm_red, want_quad;   // sm_z: [2] normals, sm_a0: [2] gathered batch, sm_y: [2]
  float* th_out; float* thT_out;         // buffer cur ^ (n_steps & 1)
};

__global__ void __launch_bounds__(kSsThreads, 1) sgld_small_kernel(const __grid_constant__ SgParams p, const __grid_constant__ SsParams q) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kSsThreads / 32;
  float* s_th = sm;                        // [sP] network (pitch lp) | [hd] log-precisions
  float* s_v = sm + q.sm_v;                // same shape
  float* s_zz = sm + q.sm_z;               // [2] the step's normals, same shape (by step parity)
  float* s_a0 = sm + q.sm_a0;              // [2][Bp][lp0] gathered batch (by step parity)
  float* s_act = sm + q.sm_act;            // inputs of Linear 1..
  float* s_ga = sm + q.sm_g0;
  float* s_gb = sm + q.sm_g1;
  float* s_yy = sm + q.sm_y;               // [2][Bp][nout] targets (by step parity), [Bp][nout] predictions
  double* s_red = reinterpret_cast<double*>(sm + q.sm_red);   // [nout + 1][warps] partial sums: r^2 per output | prior quad
  double* s_acc = s_red + (p.nout + 1) * kWarps;              // [nout + 1]
  float* s_f = reinterpret_cast<float*>(s_acc + p.nout + 1);  // LambdaLR factor of the step
  const int L = p.L, B = p.B, Bp = p.Bp, gp = q.gp, nout = p.nout;
  float* s_pred = s_yy + 2 * Bp * nout;
  const float* th_in = p.theta[p.cur];
  const bool debug = p.grad_out != nullptr;
  constexpr int kCW = kSsCompute / 32;     // compute warps
  const int zlen = q.sP + q.hd, a0len = Bp * q.sy[0].lp;

  // What a step needs that does not depend on the chain state -- the batch rows (two dependent global loads per row) and the
  // Philox normals of every parameter -- is produced ONE STEP AHEAD by the two prefetch warps while the compute warps run the
  // current step (the first step's by everybody in the prologue); double-buffered by step parity.
  auto prefetch = [&](int it2, int first, int stride) {
    const long long boff2 = p.perm_off + (long long)it2 * B;
    const long long left2 = p.num_train - boff2;
    const int rows2 = left2 < B ? (int)left2 : B;
    const int lp0 = q.sy[0].lp;
    float* a0 = s_a0 + (it2 & 1) * a0len;
    float* yb = s_yy + (it2 & 1) * Bp * nout;
    for (int i = first; i < Bp * lp0; i += stride) {                   // gather (DataLoader batch, :80,110): rows (x | 1 | 0)
      const int b = i / lp0, k = i - b * lp0;
      float v = 0.0f;
      if (b < rows2) {
        if (k < p.D) v = __ldg(p.X + __ldg(p.perm + boff2 + b) * p.D + k);
        else if (k == p.D) v = 1.0f;
      }
      a0[i] = v;
    }
    for (int i = first; i < Bp * nout; i += stride) {
      const int b = i / nout;
      yb[i] = b < rows2 ? __ldg(p.y + __ldg(p.perm + boff2 + b) * nout + (i - b * nout)) : 0.0f;
    }
    if (debug) return;
    float* zb = s_zz + (it2 & 1) * zlen;
    const uint32_t sub = (uint32_t)(p.t0 + it2);
    for (int l = 0; l < L; ++l) {
      const SgLayer& Y = p.ly[l];
      const int q4 = Y.ld >> 2;
      for (int i = first; i < Y.out * q4; i += stride) {
        const int j = i / q4, k4 = i - j * q4;
        float z[4];
        philox_normal4(p.seed, sub, STREAM_SGLD, (uint64_t)((Y.off + (long long)j * Y.ld + k4 * 4) >> 2), z);
        *reinterpret_cast<float4*>(zb + q.sy[l].soff + j * q.sy[l].lp + k4 * 4) = make_float4(z[0], z[1], z[2], z[3]);
      }
    }
    for (int i = first; i < (q.hd >> 2); i += stride) {
      float z[4];
      philox_normal4(p.seed, sub, STREAM_SGLD, (uint64_t)((p.n_head + (long long)i * 4) >> 2), z);
      *reinterpret_cast<float4*>(zb + q.sP + i * 4) = make_float4(z[0], z[1], z[2], z[3]);
    }
  };

  for (int i = tid; i < q.sm_z; i += kSsThreads) sm[i] = 0.0f;
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const SgLayer& Y = p.ly[l];
    for (int i = tid; i < Y.out * Y.ld; i += kSsThreads) {
      const int j = i / Y.ld, k = i - j * Y.ld;
      const int d = q.sy[l].soff + j * q.sy[l].lp + k;
      s_th[d] = th_in[Y.off + i];
      s_v[d] = p.V[Y.off + i];
    }
  }
  for (int i = tid; i < q.hd; i += kSsThreads) { s_th[q.sP + i] = th_in[p.n_head + i]; s_v[q.sP + i] = p.V[p.n_head + i]; }
  prefetch(0, tid, kSsThreads);
  __syncthreads();

#pragma unroll 1
  for (int it = 0; it < p.n_steps; ++it) {
    const long long t = p.t0 + it;
    const long long boff = p.perm_off + (long long)it * B;
    const long long left = p.num_train - boff;
    const int rows = left < B ? (int)left : B;
    const float scale = (float)((double)p.num_train / (double)rows);   // N / |b|  (:83)
    const bool want_quad = q.want_quad && it + 1 == p.n_steps;
    if (tid >= kSsCompute) {
      if (it + 1 < p.n_steps) prefetch(it + 1, tid - kSsCompute, kSsPrefetch);
      __syncthreads();                     // end of the step (the compute warps' matching barrier is at the bottom of the loop)
      continue;
    }
    const float* s_y = s_yy + (it & 1) * Bp * nout;
    const float* s_z = s_zz + (it & 1) * zlen;
    if (tid == 0) *s_f = it == 0 ? p.f0 : (float)pow((double)(1 + t), -0.33);   // LambdaLR (:101)
    // (first use of *s_f is behind the first barrier of the forward)
    // ---- forward: thread = 4 rows x units (j, j + half)
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const SgLayer& Y = p.ly[l];
      const SsLayer& S = q.sy[l];
      const float* a = l == 0 ? s_a0 + (it & 1) * a0len : s_act + S.a_off;
      const bool last = l + 1 == L;
      const int n_item = (Bp >> 2) * S.half;
      for (int i = tid; i < n_item; i += kSsCompute) {
        const int bq = i / S.half, j0 = i - bq * S.half, j1 = j0 + S.half;
        const bool two = j1 < Y.out;
        const float* ar = a + (bq * 4) * S.lp;
        const float* w0 = s_th + S.soff + j0 * S.lp;
        const float* w1 = two ? w0 + S.half * S.lp : w0;
        float o[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) { o[r][0] = 0.0f; o[r][1] = 0.0f; }
        for (int k = 0; k < S.lp; k += 4) {
          const float4 wa = *reinterpret_cast<const float4*>(w0 + k);
          const float4 wb = *reinterpret_cast<const float4*>(w1 + k);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float4 x = *reinterpret_cast<const float4*>(ar + r * S.lp + k);
            o[r][0] = fmaf(x.x, wa.x, o[r][0]); o[r][0] = fmaf(x.y, wa.y, o[r][0]);
            o[r][0] = fmaf(x.z, wa.z, o[r][0]); o[r][0] = fmaf(x.w, wa.w, o[r][0]);
            o[r][1] = fmaf(x.x, wb.x, o[r][1]); o[r][1] = fmaf(x.y, wb.y, o[r][1]);
            o[r][1] = fmaf(x.z, wb.z, o[r][1]); o[r][1] = fmaf(x.w, wb.w, o[r][1]);
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int b = bq * 4 + r;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c && !two) continue;
            const int j = c ? j1 : j0;
            const float v = b < rows ? o[r][c] : 0.0f;
            if (last) s_pred[b * nout + j] = v;
            else s_act[q.sy[l + 1].a_off + b * q.sy[l + 1].lp + j] = b < rows ? act_fwd(v, p.act) : 0.0f;
          }
        }
      }
      if (!last) {          // ones column and zero pad of the next layer's augmented input
        const int nin = p.ly[l + 1].in, nlp = q.sy[l + 1].lp, padw = nlp - nin;
        float* an = s_act + q.sy[l + 1].a_off;
        for (int i = tid; i < Bp * padw; i += kSsCompute) {
          const int b = i / padw, k = nin + (i - b * padw);
          an[b * nlp + k] = (k == nin && b < rows) ? 1.0f : 0.0f;
        }
      }
      named_bar_sync(1, kSsCompute);
    }
    // ---- Gaussian head (:69-74,83): r = f - y, dU/d(out) = (N / |b|) prec r, sum r^2 per output
    for (int o = 0; o < nout; ++o) {
      const float prec = expf(s_th[q.sP + o]);
      double part = 0.0;
      for (int b = tid; b < Bp; b += kSsCompute) {
        float dzv = 0.0f;
        if (b < rows) {
          const float r = s_pred[b * nout + o] - s_y[b * nout + o];
          dzv = scale * prec * r;
          part += (double)r * (double)r;
        }
        s_ga[o * gp + b] = dzv;
      }
#pragma unroll
      for (int w = 16; w > 0; w >>= 1) part += __shfl_xor_sync(0xffffffffu, part, w);
      if (lane == 0) s_red[o * kWarps + warp] = part;
    }
    named_bar_sync(1, kSsCompute);
    if (tid < nout) {
      double s = 0.0;
      for (int w = 0; w < kCW; ++w) s += s_red[tid * kWarps + w];
      s_acc[tid] = s;
    }
    if (tid == 32) s_acc[nout] = 0.0;
    // ---- backward, layer by layer from the top
    const float lr_w = p.lr_w * *s_f;
    float* g0 = s_ga;
    float* g1 = s_gb;
#pragma unroll 1
    for (int l = L - 1; l >= 0; --l) {
      const SgLayer& Y = p.ly[l];
      const SsLayer& S = q.sy[l];
      const float* a = l == 0 ? s_a0 + (it & 1) * a0len : s_act + S.a_off;
      // dz_{l-1} = (dz_l theta_l) * act'(h_l): 4 rows x 4 columns per thread, the sum over j split over 4 lanes (bits 3-4)
      if (l > 0) {
        const int k4n = (Y.in + 3) >> 2, k8 = S.k4p >> 3;
        const int n_item = 32 * k8 * (Bp >> 2);
        for (int i = tid; i < n_item; i += kSsCompute) {
          const int part = (i >> 3) & 3, r = i >> 5;
          const int k4 = (i & 7) + 8 * (r % k8), bq = r / k8;
          const bool live = k4 < k4n;
          float g[4][4];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int c = 0; c < 4; ++c) g[rr][c] = 0.0f;
          if (live) {
            const float* wp = s_th + S.soff + k4 * 4;
            const float* gq = g0 + bq * 4;
            for (int j = part; j < Y.out; j += 4) {
              const float4 w = *reinterpret_cast<const float4*>(wp + j * S.lp);
              const float4 gj = *reinterpret_cast<const float4*>(gq + j * gp);
              const float gs[4] = {gj.x, gj.y, gj.z, gj.w};
#pragma unroll
              for (int rr = 0; rr < 4; ++rr) {
                g[rr][0] = fmaf(gs[rr], w.x, g[rr][0]); g[rr][1] = fmaf(gs[rr], w.y, g[rr][1]);
                g[rr][2] = fmaf(gs[rr], w.z, g[rr][2]); g[rr][3] = fmaf(gs[rr], w.w, g[rr][3]);
              }
            }
          }
          float mine[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              float v = g[rr][c];
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              if (c == part) mine[rr] = v;
            }
          }
          const int k = k4 * 4 + part;
          if (live && k < Y.in) {
            float out4[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) out4[rr] = mine[rr] * act_bwd(a[(bq * 4 + rr) * S.lp + k], p.act);
            *reinterpret_cast<float4*>(g1 + k * gp + bq * 4) = make_float4(out4[0], out4[1], out4[2], out4[3]);
          }
        }
      }
      named_bar_sync(1, kSsCompute);
      // dW tile (units (j, j + half) x 4 columns) + prior gradient W / std^2 (:65-66) + pSGLD update of those entries
      {
        const int q4 = S.lp >> 2;
        const int n_item = S.half * q4;
        double quad = 0.0;
        for (int i = tid; i < n_item; i += kSsCompute) {
          const int j0 = i / q4, k = (i - j0 * q4) * 4, j1 = j0 + S.half;
          if (k >= Y.ld) continue;
          const bool two = j1 < Y.out;
          const float* ga = g0 + j0 * gp;
          const float* gb = two ? g0 + j1 * gp : ga;
          float d[2][4];
#pragma unroll
          for (int c = 0; c < 4; ++c) { d[0][c] = 0.0f; d[1][c] = 0.0f; }
          for (int b = 0; b < Bp; b += 4) {
            const float4 u = *reinterpret_cast<const float4*>(ga + b);
            const float4 v = *reinterpret_cast<const float4*>(gb + b);
            const float us[4] = {u.x, u.y, u.z, u.w}, vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const float4 x = *reinterpret_cast<const float4*>(a + (b + rr) * S.lp + k);
              d[0][0] = fmaf(us[rr], x.x, d[0][0]); d[0][1] = fmaf(us[rr], x.y, d[0][1]);
              d[0][2] = fmaf(us[rr], x.z, d[0][2]); d[0][3] = fmaf(us[rr], x.w, d[0][3]);
              d[1][0] = fmaf(vs[rr], x.x, d[1][0]); d[1][1] = fmaf(vs[rr], x.y, d[1][1]);
              d[1][2] = fmaf(vs[rr], x.z, d[1][2]); d[1][3] = fmaf(vs[rr], x.w, d[1][3]);
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h && !two) continue;
            const int j = h ? j1 : j0;
            const int w0 = S.soff + j * S.lp + k;
            const long long e = Y.off + (long long)j * Y.ld + k;          // multiple of 4: one Philox block per quad
            float tt[4], g[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              tt[c] = s_th[w0 + c];
              g[c] = d[h][c];
              if (k + c < Y.in) {
                g[c] = fmaf(Y.inv_var, tt[c], g[c]);
                if (want_quad) quad += (double)(Y.inv_var * tt[c]) * (double)tt[c];
              } else if (k + c > Y.in) g[c] = 0.0f;                          // layout padding: no gradient, no noise, stays zero
            }
            if (debug) {
              *reinterpret_cast<float4*>(p.grad_out + e) = make_float4(g[0], g[1], g[2], g[3]);
              continue;
            }
            const float4 z4 = *reinterpret_cast<const float4*>(s_z + w0);     // Philox (seed, t), block e / 4: drawn a step ahead
            const float z[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (k + c <= Y.in) {
                float vv = s_v[w0 + c];
                psgld_update1(tt[c], vv, g[c], z[c], lr_w, p.alpha, p.lam, p.precond_mode);
                s_th[w0 + c] = tt[c];
                s_v[w0 + c] = vv;
              }
            }
          }
        }
        if (want_quad) {
#pragma unroll
          for (int w = 16; w > 0; w >>= 1) quad += __shfl_xor_sync(0xffffffffu, quad, w);
          if (lane == 0) s_red[nout * kWarps + warp] = quad;
        }
      }
      named_bar_sync(1, kSsCompute);
      if (want_quad && tid == 0) {
        double s = s_acc[nout];
        for (int w = 0; w < kCW; ++w) s += s_red[nout * kWarps + w];
        s_acc[nout] = s;
      }
      float* tsw = g0; g0 = g1; g1 = tsw;
    }
    // ---- U of the launch's last step (evaluated before the update of the log-precisions)
    if (want_quad && tid == 0) {
      float total_ll = 0.0f, total_lp = 0.0f;
      for (int o = 0; o < nout; ++o) {
        const float lp = s_th[q.sP + o], prec = expf(lp);
        const float s2 = (float)s_acc[o];
        total_ll += -0.5f * prec * s2 + (float)rows * (0.5f * lp - 0.91893853320467274178f);
        total_lp += p.al * p.log_be + (p.al - 1.0f) * lp - p.be * prec - p.lgamma_al + lp;
      }
      const float logp = -0.5f * (float)s_acc[nout] - p.log_const + total_lp;
      p.loss_out[0] = -(total_ll * scale + logp);
    }
    named_bar_sync(1, kSsCompute);
    // ---- log-precision group (:62, :96-98): closed-form gradient, same update rule, lr_noise
    if (tid < (q.hd >> 2)) {
      const float lr_n = p.lr_n * *s_f;
      const long long e = p.n_head + (long long)tid * 4;
      float tt[4], vv[4], g[4], z[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tt[c] = s_th[q.sP + tid * 4 + c];
        vv[c] = s_v[q.sP + tid * 4 + c];
        g[c] = 0.0f;
        const int o = tid * 4 + c;
        if (o < nout) {
          const float prec = expf(tt[c]);
          const float s2 = (float)s_acc[o];
          g[c] = -scale * (-0.5f * prec * s2 + 0.5f * (float)rows) - (p.al - p.be * prec);
        }
      }
      if (debug) {
#pragma unroll
        for (int c = 0; c < 4; ++c) p.grad_out[e + c] = g[c];
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) z[c] = s_z[q.sP + tid * 4 + c];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (tid * 4 + c < nout) psgld_update1(tt[c], vv[c], g[c], z[c], lr_n, p.alpha, p.lam, p.precond_mode);
          else tt[c] = 0.0f;
          s_th[q.sP + tid * 4 + c] = tt[c];
          s_v[q.sP + tid * 4 + c] = vv[c];
        }
      }
    }
    __syncthreads();
  }
  if (debug) return;
  // ---- the chain state goes back to global memory: theta / thetaT of buffer cur ^ (n_steps & 1), V
  for (int l = 0; l < L; ++l) {
    const SgLayer& Y = p.ly[l];
    for (int i = tid; i < Y.out * Y.ld; i += kSsThreads) {
      const int j = i / Y.ld, k = i - j * Y.ld;
      const int d = q.sy[l].soff + j * q.sy[l].lp + k;
      const float w = s_th[d];
      q.th_out[Y.off + i] = w;
      p.V[Y.off + i] = s_v[d];
      if (l > 0 && k < Y.in) q.thT_out[Y.offT + (long long)k * Y.ldT + j] = w;
    }
  }
  for (int i = tid; i < q.hd; i += kSsThreads) { q.th_out[p.n_head + i] = s_th[q.sP + i]; p.V[p.n_head + i] = s_v[q.sP + i]; }
}

// 0 = the single-SM kernel takes this chain; 1 = it does not fit
static int ss_plan(const SgParams& p, SsParams* q, size_t* smem) {
  memset(q, 0, sizeof(*q));
  q->gp = p.Bp + 4 - (p.Bp & 4);                                   // gp / 4 odd
  q->hd = (int)(p.n_theta - p.n_head);
  long long sP = 0, act = 0;
  int max_w = p.nout;
  for (int l = 0; l < p.L; ++l) {
    const SgLayer& Y = p.ly[l];
    SsLayer& S = q->sy[l];
    S.lp = ((Y.ld >> 2) & 1) ? Y.ld : Y.ld + 4;
    S.half = (Y.out + 1) >> 1;
    S.k4p = (((Y.in + 3) >> 2) + 7) & ~7;
    S.soff = (int)sP; sP += (long long)Y.out * S.lp;
    S.a_off = (int)act; act += (long long)p.Bp * S.lp;
    if (Y.out > max_w) max_w = Y.out;
    if (Y.in + 1 > max_w) max_w = Y.in + 1;
  }
  q->sP = (int)sP; q->max_w = max_w;
  long long f = sP + q->hd;
  q->sm_v = (int)f; f += sP + q->hd;
  q->sm_z = (int)f; f += 2 * (sP + q->hd);
  q->sm_a0 = (int)f; f += 2LL * p.Bp * q->sy[0].lp;
  q->sm_act = (int)f; f += act;
  q->sm_g0 = (int)f; f += (long long)max_w * q->gp;
  q->sm_g1 = (int)f; f += (long long)max_w * q->gp;
  q->sm_y = (int)f; f += (3LL * p.Bp * p.nout + 3) & ~3LL;
  q->sm_red = (int)f; f += 2 * ((long long)(p.nout + 1) * (kSsThreads / 32) + p.nout + 1) + 4;
  *smem = (size_t)f * 4;
  return *smem <= (size_t)227 * 1024 ? 0 : 1;
}

static long long* g_sg_trace = nullptr;   // debug: set by bnn_sgld_trace
static int g_sg_grid_cap = 0;              // bnn_sgld_set_grid_cap
static int g_sg_trace_grid = 0;

extern "C" int32_t bnn_sgld_set_grid_cap(int32_t ctas) {
  if (ctas < 0) BNN_FAIL("bnn_sgld_set_grid_cap: ctas must be >= 0 (0 = no cap)");
  g_sg_grid_cap = ctas;
  return 0;
}

extern "C" int32_t bnn_sgld_run(const BnnSgldChain* c, int32_t cur, int64_t t0, int64_t perm_off, int32_t n_steps, float* loss_out,
                                float* grad_out, void* stream) {
  if (!c || (cur != 0 && cur != 1) || n_steps < 1 || t0 < 0 || perm_off < 0) BNN_FAIL("bnn_sgld_run: bad arguments");
  if (!c->X || !c->y || !c->perm || c->num_train < 1) BNN_FAIL("bnn_sgld_run: X, y, perm and num_train are required");
  if (grad_out && n_steps != 1) BNN_FAIL("bnn_sgld_run: the gradient-only mode runs exactly one step");
  if (perm_off + (int64_t)(n_steps - 1) * c->batch >= c->num_train)
    BNN_FAIL("bnn_sgld_run: %d steps of batch %d from offset %lld run past the epoch (%lld rows)", n_steps, c->batch,
             (long long)perm_off, (long long)c->num_train);
  SgPlan plan;
  if (sg_plan(c, &plan, true, loss_out != nullptr)) return -1;
  SgParams& p = plan.p;
  if ((long long)n_steps * p.jobs_per_step >= (1LL << 31)) BNN_FAIL("bnn_sgld_run: too many steps in one launch (%d x %d jobs)", n_steps, p.jobs_per_step);
  p.f0 = (float)pow((double)(1 + t0), -0.33);                  // LambdaLR factor of the first step; later ones are computed in-kernel
  p.X = c->X; p.y = c->y; p.perm = (const long long*)c->perm; p.num_train = c->num_train; p.perm_off = perm_off;
  p.t0 = t0; p.n_steps = n_steps; p.cur = cur;
  p.loss_out = loss_out; p.grad_out = grad_out;
  cudaStream_t st = (cudaStream_t)stream;
  {
    // networks that fit one SM: the single-CTA kernel (chain state in shared memory for the whole launch)
    SsParams q;
    size_t ssmem = 0;
    // ... and whose step is small enough that one SM beats the hand-overs of the dataflow kernel (measured: 13-50-1 at batch 32
    // 10.0 vs 15.9 us / step, at batch 128 18.4 vs 16.8, 9-100-100-1 at batch 32 26.7 vs 26.6): packed parameters x batch rows <=
    // kSsMaxWork.
    // BNN_SGLD_SMALL = 0 never, 1 default rule, 2 whenever it fits (tests, probes).
    const char* e = getenv("BNN_SGLD_SMALL");
    const int mode = e ? atoi(e) : 1;
    const bool small_enough = mode == 2 || p.n_head * (long long)p.Bp <= kSsMaxWork;
    if (mode != 0 && small_enough && ss_plan(p, &q, &ssmem) == 0) {
      const int nb = cur ^ (n_steps & 1);
      q.th_out = p.theta[nb]; q.thT_out = p.thetaT[nb];
      q.want_quad = loss_out != nullptr;
      BNN_CUDA(cudaFuncSetAttribute(sgld_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
      sgld_small_kernel<<<1, kSsThreads, ssmem, st>>>(p, q);
      BNN_CUDA(cudaGetLastError());
      return 0;
    }
  }
  // accumulators, job / completion counters and error flag start at zero for every launch
  BNN_CUDA(cudaMemsetAsync(p.acc, 0, (size_t)plan.tail_bytes, st));
  BNN_CUDA(cudaFuncSetAttribute(sgld_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  int per_sm = 0;
  BNN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sgld_step_kernel, kSgBlock, plan.smem));
  if (per_sm < 1) BNN_FAIL("bnn_sgld_run: the step kernel does not fit on an SM");
  int grid = plan.grid;
  if (grid > per_sm * sm_count()) grid = per_sm * sm_count();
  // bnn_sgld_set_grid_cap / BNN_SGLD_GRID cap the CTAs of one chain: several chains launched on different streams then share the GPU (each launch is
  // cooperative, so a chain only starts when all of its CTAs fit next to the others')
  if (g_sg_grid_cap >= 1 && g_sg_grid_cap < grid) grid = g_sg_grid_cap;
  if (const char* e = getenv("BNN_SGLD_GRID")) { const int v = atoi(e); if (v >= 1 && v < grid) grid = v; }
  p.trace = g_sg_trace;
  if (const char* e = getenv("BNN_SGLD_DBG")) p.dbg = atoi(e);
  g_sg_trace_grid = grid;
  if (p.trace) BNN_CUDA(cudaMemsetAsync(p.trace, 0, (size_t)grid * kSgTraceRecs * kSgTraceW * sizeof(long long), st));
  {
    const size_t gsmem = (size_t)32 * (p.ly[0].ld + 1) * 4;
    BNN_CUDA(cudaFuncSetAttribute(sgld_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
    sgld_gather_kernel<<<(unsigned)((p.B + FT_M - 1) / FT_M), kSgThreads, gsmem, st>>>(p);
    BNN_CUDA(cudaGetLastError());
  }
  void* args[] = {(void*)&p};
  BNN_CUDA(cudaLaunchCooperativeKernel((void*)sgld_step_kernel, dim3((unsigned)grid), dim3(kSgBlock), args, plan.smem, st));
  return 0;
}

extern "C" int32_t bnn_sgld_perm(uint64_t seed, uint32_t epoch, int64_t N, int64_t offset, int64_t count, int64_t* out, void* stream) {
  if (!out || N < 1 || offset < 0 || count < 1 || offset + count > N) BNN_FAIL("bnn_sgld_perm: bad arguments");
  int bits = 1;
  while ((1ULL << bits) < (unsigned long long)N) ++bits;
  const int hb = (bits + 1) / 2;
  sgld_perm_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, epoch, N, offset, count, hb, (long long*)out);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

// debug: enable = 1 allocates the per-CTA job timeline written by the following bnn_sgld_run launches (jobs of the last two
// steps of each launch); with host_out != NULL copies it out: [grid][recs][6] = {it << 32 | type << 24 | layer << 16 | tile,
// clock64 at grab, at inputs-ready, at done, globaltimer (ns) at grab, at done}; returns grid in *grid_out and recs in *recs_out
extern "C" int32_t bnn_sgld_trace(int32_t enable, long long* host_out, int32_t* grid_out, int32_t* recs_out) {
  const size_t bytes = (size_t)1024 * kSgTraceRecs * kSgTraceW * sizeof(long long);
  if (enable && !g_sg_trace) {
    BNN_CUDA(cudaMalloc(&g_sg_trace, bytes));
    BNN_CUDA(cudaMemset(g_sg_trace, 0, bytes));
  }
  if (host_out && g_sg_trace) {
    BNN_CUDA(cudaDeviceSynchronize());
    BNN_CUDA(cudaMemcpy(host_out, g_sg_trace, (size_t)g_sg_trace_grid * kSgTraceRecs * kSgTraceW * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  if (grid_out) *grid_out = g_sg_trace_grid;
  if (recs_out) *recs_out = kSgTraceRecs;
  if (!enable && g_sg_trace) { cudaFree(g_sg_trace); g_sg_trace = nullptr; }
  return 0;
}

// 0 = ok; 1 = a dependency wait of the last launches timed out (never expected; the chain state is then undefined)
extern "C" int32_t bnn_sgld_status(const BnnSgldChain* c, int32_t* status_host, void* stream) {
  if (!c || !status_host) BNN_FAIL("bnn_sgld_status: bad arguments");
  SgPlan plan;
  if (sg_plan(c, &plan, true)) return -1;
  BNN_CUDA(cudaMemcpyAsync(status_host, plan.p.err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  BNN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
