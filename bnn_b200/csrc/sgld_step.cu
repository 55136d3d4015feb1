// Row f1 of SURVEY.md 8: the whole SGLD minibatch step of BNN_SGDMC.sgld_steps (BNN_SGDMC.py:76-91) as ONE persistent
// cooperative kernel that runs n_steps steps per launch:
//   minibatch gather (DataLoader batch, :80,110) -> L forward layers (util.py:31-33) -> Gaussian log-likelihood + priors and
//   dU/d(out) (:61-74,83) -> backward (dW, dh; what loss.backward() computes, :85) -> preconditioned SGLD update with
//   in-kernel Philox noise and the LambdaLR factor (:86-87,96-101).
// No library GEMM, no elementwise launches: every phase is a set of tile jobs spread over the CTAs, phases are separated by
// a grid barrier (2 L barriers per step).  The gradient never exists in memory: the CTA that finishes a dW tile applies the
// update to that tile of (theta, V) in its epilogue.
//
// Why FP32 FFMA and not tcgen05 here (north_star: "tensor cores only where width x S makes it a real contraction"): at
// cfg 5 a step is 438 MFLOP spread over 8 dependent phases; one phase is a [128 x 512 x 512] product = 0.26 MFLOP per
// CTA-tile when spread over 128 SMs, i.e. ~2k cycles of FFMA, the same order as the L2 -> SM ingest of the tile's operands
// and as the grid barrier.  The step is latency / barrier bound, not FLOP bound; FP32 also keeps the gradient bit-comparable
// (1e-6) with the reference's autograd.
//
// Layout: every product is brought into the "NT" form C[m, n] = sum_k A[m, k] * B[n, k] with both operands K-contiguous:
//   forward  l : C = z_l[b, j]        A = h_l[b, 0:ld_l]   (augmented (h | 1 | 0))   B = theta_l[j, 0:ld_l]  ((W | b | 0))
//   dW       l : C = dW_l[j, k]       A = dzT_l[j, 0:Bp]                             B = hT_l[k, 0:Bp]       (row in_l = ones -> db)
//   dh       l : C = dh_l[b, k]       A = dz_l[b, 0:ldz_l]                           B = thetaT_l[k, 0:ldT_l]
// so activations and output gradients are stored twice (row-major and transposed) by the producing epilogue, and the
// weights of layers >= 1 carry a transposed copy maintained by the update epilogue.  theta / thetaT are double-buffered by
// step parity: all reads of a step see theta_cur, all updates write theta_next -- no intra-step hazards, no extra barrier.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "philox.cuh"

namespace bnn {

constexpr int kSgThreads = 256;
constexpr int kSgKC = 64;            // K chunk (floats)
constexpr int kSgLds = kSgKC + 4;    // smem row pitch: 68 words = 4 (mod 32) -> 8 consecutive rows hit 32 distinct banks
// Stage counts per tile family.  A tile's operands are tiny (<= 110 KB) but every access is an L2 round trip of ~450
// cycles (DRAM ~1000), so the ring is deep enough to put the WHOLE K extent of a tile in flight at once (K <= 768 resp.
// 384): tile time = one latency + transfer, instead of one latency per two chunks.
constexpr int kSgStagesF = 12;   // 32 x 16 tiles: 12 x 13,056 B
constexpr int kSgStagesW = 6;    // 64 x 32 tiles:  6 x 26,112 B
constexpr int kSgMaxStages = 12;

struct SgLayer {
  int in, out, ld, ldz, ldT;
  long long off, offT;       // float offsets into theta / thetaT
  float* h;                  // input of Linear l, augmented: [Bp, ld]   (null for l = 0: gathered from X)
  float* hT;                 // [ld, Bp]                                  (null for l = 0)
  float* dz;                 // dU/d(pre-activation output of Linear l): [Bp, ldz]
  float* dzT;                // [out, Bp]
  float inv_var;             // 1 / std_l^2 of the weight prior, std_l = gain * sqrt(2 / (in + out))   (BNN_SGDMC.py:65)
};

struct SgParams {
  int L, act, B, Bp, nout, D;
  SgLayer ly[BNN_MAX_LINEAR];
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
  double* acc;                         // [nout] sum r^2 per output | [nout] prior quad (one slot used) -- zero before launch
  unsigned int* bar;                   // grid barrier counter, zero before launch
  int* err;
  float* loss_out;                     // U of the LAST step of the launch (may be null)
  float* grad_out;                     // debug: when non-null the step writes dU/dtheta here and does NOT update
  int dbg;                             // debug (BNN_SGLD_DBG): 1 skip the FMA loop, 2 skip the dense loads, 4 skip the epilogues
  long long* trace;                    // debug: [grid][2 * L phases][2] clock64 (work done, barrier passed) of the LAST step
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

// Grid barrier on a monotonically increasing counter.  Co-residency of all CTAs is guaranteed by the cooperative launch.
// A spin limit turns a (never expected) deadlock into an error code instead of a hung GPU.
__device__ __forceinline__ bool grid_barrier(unsigned int* counter, unsigned int& target, int* err, int* s_flag) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    // release at gpu scope: the CTA's writes (ordered before this by the bar.sync above) are visible to whoever acquires
    // the counter; cheaper than __threadfence() (= fence.sc.gpu) followed by a relaxed add
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned long long spins = 0;
    int bad = 0;
    while ((int)(ld_acquire_u32(counter) - target) < 0) {
      if (++spins > (1ull << 26)) { atomicExch(err, 1); bad = 1; break; }
      if ((spins & 1023) == 0 && *((volatile int*)err) != 0) { bad = 1; break; }
    }
    *s_flag = bad;
  }
  __syncthreads();
  return *s_flag == 0;
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
enum { OP_DENSE = 0, OP_GATHER_X = 1, OP_GATHER_XT = 2 };
struct SgOperand {
  int kind;
  const float* p;      // DENSE: [rows, ld]
  int ld;
  int rows_valid;      // rows >= rows_valid read as zero
  int k_valid;         // k >= k_valid reads as zero (multiple of 4)
};
// what the gather kinds need (X[perm[b] * D + k]; batch rows >= rows_live are zero: short last batch of an epoch)
struct SgGather {
  const float* __restrict__ X;
  const long long* perm;   // this step's batch rows (shared-memory copy of perm[off .. off + B))
  int D, B, rows_live;
};

// gathered operands (synchronous): rows = batch index b, K = input feature (forward of Linear 0: kT = false) or rows = input
// feature, K = batch index (dW of Linear 0: kT = true); augmented with the ones column / row at feature == D.  Four
// independent loads are in flight per thread before the first store.
template <int ROWS, bool kT>
__device__ __forceinline__ void load_gather(const SgGather& gq, int row0, int k0, float* dst) {
#pragma unroll 1
  for (int base = threadIdx.x; base < ROWS * kSgKC; base += 4 * kSgThreads) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * kSgThreads;
      const int r = kT ? idx % ROWS : idx >> 6, kk = kT ? idx / ROWS : idx & 63;
      const int b = kT ? k0 + kk : row0 + r, f = kT ? row0 + r : k0 + kk;
      v[u] = 0.0f;
      if (idx < ROWS * kSgKC && b < gq.rows_live) {
        if (f < gq.D) v[u] = __ldg(gq.X + gq.perm[b] * gq.D + f);
        else if (f == gq.D) v[u] = 1.0f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + u * kSgThreads;
      if (idx < ROWS * kSgKC) dst[(kT ? idx % ROWS : idx >> 6) * kSgLds + (kT ? idx / ROWS : idx & 63)] = v[u];
    }
  }
}

// ---- tile core: red[TM x TN] = sum_k A[m0 + ., k] B[n0 + ., k], left in shared memory (row-major, after the K-slices are summed)
// Operands: 16-byte cp.async into padded shared-memory rows, completion per ring stage on an mbarrier (every thread posts a
// deferred arrival that fires when its own copies of the stage have landed: no per-chunk __syncthreads).  Rows beyond the
// operand are not copied (garbage rows only feed output rows / columns the epilogues discard); the K tail is skipped.
// Compute: FP32 FFMA, thread tile 4 x 4 (rows ty + TY * i, columns tx + TX * j); the 256 threads form KS = 256 / (TY * TX)
// K-slices, slice ks takes the float4 groups g = ks, ks + KS, ... of every chunk.
// Measured alternatives (profiles/r02_sgld_step_notes.md): mma.sync m16n8k8 TF32 with 3xTF32 compensation is NOT faster --
// the legacy warp-level tensor path of this chip issues one TF32 mma per ~32 cycles per SM sub-partition, i.e. the FFMA
// rate, and the compensation needs three of them; one cp.async.bulk per 256-byte operand row was slower than 16-byte
// cp.async (432 small bulk copies per tile); 16 warps instead of 8 did not change the K loop.
template <int TM, int TN, int NS>
__device__ __forceinline__ void tile_core(const SgOperand& A, const SgOperand& Bop, const SgGather& gq, int m0, int n0, int K,
                                          float* smem, uint64_t* bars, uint32_t& phase_bits, int dbg, long long* probe) {
  constexpr int TY = TM / 4, TX = TN / 4, THR = TY * TX, KS = kSgThreads / THR;
  constexpr int STAGE = (TM + TN) * kSgLds;
  static_assert(KS >= 1 && KS * THR == kSgThreads && NS <= kSgMaxStages, "tile shape");
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = (tid / TX) % TY, ks = tid / THR;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  if (probe && tid == 0) probe[0] = clock64();
  const int nchunks = (K + kSgKC - 1) / kSgKC;
  const bool a_dense = A.kind == OP_DENSE, b_dense = Bop.kind == OP_DENSE;
  const int va = min(TM, A.rows_valid - m0), vb = min(TN, Bop.rows_valid - n0);   // valid rows of the tile (>= 1)
  auto issue_dense = [&](int c) {
    if (dbg & 2) return;
    const int k0 = c * kSgKC;
    const int kq = (min(kSgKC, K - k0) + 3) >> 2;                  // float4s per row in this chunk
    float* st = smem + (c % NS) * STAGE;
    const uint32_t d0 = smem_u32(st);
#pragma unroll
    for (int idx = tid; idx < (TM + TN) * (kSgKC / 4); idx += kSgThreads) {
      const int r = idx >> 4, c4 = idx & 15;
      const bool isA = r < TM;
      const int rr = isA ? r : r - TM;
      const SgOperand& op = isA ? A : Bop;
      if (c4 < kq && rr < (isA ? va : vb) && (isA ? a_dense : b_dense))
        cp_async16(d0 + (uint32_t)(r * kSgLds + c4 * 4) * 4u, op.p + (long long)((isA ? m0 : n0) + rr) * op.ld + k0 + c4 * 4, 16);
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[c % NS])) : "memory");
  };
  auto issue_gather = [&](int c) {
    if (dbg & 2) return;
    float* st = smem + (c % NS) * STAGE;
    if (!a_dense) load_gather<TM, false>(gq, m0, c * kSgKC, st);
    if (!b_dense) load_gather<TN, true>(gq, n0, c * kSgKC, st + TM * kSgLds);
  };
  // prologue: everything the ring holds
  {
    const int npre = min(nchunks, NS);
#pragma unroll 1
    for (int c = 0; c < npre; ++c) issue_dense(c);
    if (!a_dense || !b_dense) {
#pragma unroll 1
      for (int c = 0; c < npre; ++c) issue_gather(c);
      __syncthreads();                     // gathered rows are plain shared-memory stores
    }
  }
  if (probe && tid == 0) probe[1] = clock64();
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    const int s = c % NS;
    if (!(dbg & 2)) mbar_wait_spin(&bars[s], (phase_bits >> s) & 1u);
    phase_bits ^= 1u << s;
    const float* As = smem + s * STAGE + ty * kSgLds;
    const float* Bs = smem + s * STAGE + (TM + tx) * kSgLds;
    const int rem = K - c * kSgKC;
    const int gmax = (dbg & 1) ? 0 : (rem >= kSgKC ? kSgKC / 4 : (rem + 3) >> 2);
#pragma unroll 1
    for (int g = ks; g < gmax; g += KS) {
      float4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(As + TY * i * kSgLds + g * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bs + TX * j * kSgLds + g * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
    if (c + NS < nchunks) {                // refill this stage (only for K beyond the ring: layers wider than 768 / 384)
      __syncthreads();                     // every warp is done with chunk c
      issue_dense(c + NS);
      if (!a_dense || !b_dense) { issue_gather(c + NS); __syncthreads(); }
    }
  }
  __syncthreads();                         // all slices done with the stages: reuse them for the cross-slice reduction
  if (probe && tid == 0) probe[2] = clock64();
  float* red = smem + TM * TN;             // [KS][TM * TN] partials behind the [TM * TN] result
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[ks * (TM * TN) + (ty + TY * i) * TN + tx + TX * j] = acc[i][j];
  __syncthreads();
#pragma unroll 1
  for (int q = tid; q < TM * TN / 4; q += kSgThreads) {
    float4 v = *reinterpret_cast<const float4*>(red + q * 4);
#pragma unroll
    for (int s2 = 1; s2 < KS; ++s2) {
      const float4 w = *reinterpret_cast<const float4*>(red + s2 * (TM * TN) + q * 4);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    *reinterpret_cast<float4*>(smem + q * 4) = v;
  }
  __syncthreads();
  if (probe && tid == 0) probe[3] = clock64();
}

// ---- update of one parameter (shared by the dW epilogue and the log-precision job) ----------------------------------------
__device__ __forceinline__ void psgld_update1(float& t, float& v, float g, float z, float lr, float alpha, float lam, int mode) {
  v = alpha * v + (1.0f - alpha) * g * g;
  const float G = mode == 0 ? 1.0f / (lam + sqrtf(v)) : rsqrtf(v + lam);
  t = t - (0.5f * lr * G * g + sqrtf(lr * G) * z);
}

constexpr int FT_M = 32, FT_N = 16;   // forward / dh tiles over [batch, features]
constexpr int WT_M = 64, WT_N = 32;   // dW tiles over [out, ld]

__global__ void __launch_bounds__(kSgThreads, 1) sgld_step_kernel(const __grid_constant__ SgParams p) {
  extern __shared__ __align__(16) float sg_smem[];
  __shared__ int s_flag;
  __shared__ long long s_perm[1024];   // this step's batch rows
  __shared__ __align__(8) uint64_t s_bars[kSgMaxStages];   // one per ring stage: "the chunk's bytes have landed"
  unsigned int bar_target = 0;
  uint32_t phase_bits = 0;             // parity to wait for next, per stage (uniform across the CTA)
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < kSgMaxStages; ++i) mbar_init(&s_bars[i], kSgThreads);   // every thread posts one (deferred) arrival per use
    fence_barrier_init();
  }
  __syncthreads();
  const int L = p.L;
  const bool debug = p.grad_out != nullptr;

#pragma unroll 1
  for (int it = 0; it < p.n_steps; ++it) {
    const long long t = p.t0 + it;
    const int cur = (p.cur + it) & 1;
    const float* th = p.theta[cur];
    const float* thT = p.thetaT[cur];
    float* th_next = p.theta[cur ^ 1];
    float* thT_next = p.thetaT[cur ^ 1];
    const long long boff = p.perm_off + (long long)it * p.B;
    const long long left = p.num_train - boff;
    const int rows = left < p.B ? (int)left : p.B;                 // the last batch of an epoch may be short (:80)
    // the batch's rows, read once per step per CTA (the gathers and the head index X / y through it); the previous step's
    // readers are behind that step's last grid barrier
    for (int b = tid; b < rows; b += kSgThreads) s_perm[b] = p.perm[boff + b];
    __syncthreads();
    SgGather gq;
    gq.X = p.X; gq.perm = s_perm; gq.D = p.D; gq.B = p.B; gq.rows_live = rows;
    const float scale = (float)((double)p.num_train / (double)rows);   // N / |b|  (:83)
    const float f = (float)pow((double)(1 + t), -0.33);            // LambdaLR: np.float32((1 + it) ** -0.33)  (:101)
    const float lr_w = p.lr_w * f, lr_n = p.lr_n * f;
    const bool want_loss = p.loss_out != nullptr && it == p.n_steps - 1;
    const int bm = (p.B + FT_M - 1) / FT_M;                         // batch-row tiles

    // phases 0 .. L-1: forward of Linear ph; phases L .. 2L-1: backward + update of Linear 2L-1-ph
#pragma unroll 1
    for (int ph = 0; ph < 2 * L; ++ph) {
      const bool fwd = ph < L;
      const int l = fwd ? ph : 2 * L - 1 - ph;
      const SgLayer& Y = p.ly[l];
      const int wm = (Y.out + WT_M - 1) / WT_M;
      const int n_dw = fwd ? 0 : wm * ((Y.ld + WT_N - 1) / WT_N);
      const int n_ft = bm * (fwd ? (Y.out + FT_N - 1) / FT_N : (l > 0 ? (Y.in + FT_N - 1) / FT_N : 0));
      const int n_jobs = n_dw + n_ft + ((!fwd && l == L - 1) ? 1 : 0);
      long long* probe = (p.trace && it == p.n_steps - 1 && blockIdx.x == 1) ? p.trace + 1024 * 2 * BNN_MAX_LINEAR * 2 + ph * 8 : nullptr;
      if (probe && tid == 0) probe[5] = clock64();
#pragma unroll 1
      for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        if (job < n_dw) {
          // ---- dW_l tile [64 x 32] over K = batch, then the pSGLD update of that tile of theta / V
          SgOperand A, Bo;
          A.kind = OP_DENSE; A.p = Y.dzT; A.ld = p.Bp; A.rows_valid = Y.out; A.k_valid = p.Bp;
          Bo.kind = l == 0 ? OP_GATHER_XT : OP_DENSE; Bo.p = Y.hT; Bo.ld = p.Bp; Bo.rows_valid = Y.ld; Bo.k_valid = p.Bp;
          const int m0 = (job % wm) * WT_M, n0 = (job / wm) * WT_N;
          tile_core<WT_M, WT_N, kSgStagesW>(A, Bo, gq, m0, n0, p.Bp, sg_smem, s_bars, phase_bits, p.dbg, job == blockIdx.x ? probe : nullptr);
          double quad_local = 0.0;
          if (!(p.dbg & 4)) {
#pragma unroll 1
            for (int q = tid; q < WT_M * WT_N / 4; q += kSgThreads) {
              const int j = m0 + (q * 4) / WT_N, k0 = n0 + (q * 4) % WT_N;
              if (j >= Y.out || k0 >= Y.ld) continue;
              const float4 v = *reinterpret_cast<const float4*>(sg_smem + q * 4);
              const long long e = Y.off + (long long)j * Y.ld + k0;          // multiple of 4: one Philox block per quad
              const float4 t4 = *reinterpret_cast<const float4*>(th + e);
              float tt[4] = {t4.x, t4.y, t4.z, t4.w};
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
              const float4 v4 = *reinterpret_cast<const float4*>(p.V + e);
              float vv[4] = {v4.x, v4.y, v4.z, v4.w};
              float z[4];
              philox_normal4(p.seed, (uint32_t)t, STREAM_SGLD, (uint64_t)(e >> 2), z);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int k = k0 + c;
                if (k <= Y.in) psgld_update1(tt[c], vv[c], g[c], z[c], lr_w, p.alpha, p.lam, p.precond_mode);
                else tt[c] = 0.0f;
              }
              *reinterpret_cast<float4*>(th_next + e) = make_float4(tt[0], tt[1], tt[2], tt[3]);
              *reinterpret_cast<float4*>(p.V + e) = make_float4(vv[0], vv[1], vv[2], vv[3]);
              if (l > 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const int k = k0 + c;
                  if (k < Y.in) thT_next[Y.offT + (long long)k * Y.ldT + j] = tt[c];
                }
              }
            }
          }
          __syncthreads();                 // the result tile in shared memory is free for the next job's loads
          if (want_loss) {
            // prior energy sum W^2 / std^2 of the step's INPUT parameters (the loss is evaluated before the update, :81-83)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) quad_local += __shfl_xor_sync(0xffffffffu, quad_local, o);
            if ((tid & 31) == 0 && quad_local != 0.0) atomicAdd(&p.acc[p.nout], quad_local);
          }
        } else if (job < n_dw + n_ft) {
          // ---- [32 x 16] tile over [batch, features]: forward of Linear l, or dh of Linear l (-> dz of Linear l - 1)
          const int jj = job - n_dw;
          const int m0 = (jj % bm) * FT_M, n0 = (jj / bm) * FT_N;
          SgOperand A, Bo;
          A.kind = OP_DENSE; Bo.kind = OP_DENSE;
          int K;
          if (fwd) {
            if (l == 0) A.kind = OP_GATHER_X;
            A.p = Y.h; A.ld = Y.ld; A.rows_valid = p.B; A.k_valid = Y.ld;
            Bo.p = th + Y.off; Bo.ld = Y.ld; Bo.rows_valid = Y.out; Bo.k_valid = Y.ld;
            K = Y.ld;
          } else {
            A.p = Y.dz; A.ld = Y.ldz; A.rows_valid = p.B; A.k_valid = Y.ldz;
            Bo.p = thT + Y.offT; Bo.ld = Y.ldT; Bo.rows_valid = Y.in; Bo.k_valid = Y.ldT;
            K = Y.ldz;
          }
          tile_core<FT_M, FT_N, kSgStagesF>(A, Bo, gq, m0, n0, K, sg_smem, s_bars, phase_bits, p.dbg, job == blockIdx.x ? probe : nullptr);
          if (!(p.dbg & 4) && tid < FT_M * FT_N / 4) {
            const int b = m0 + (tid * 4) / FT_N, c0 = n0 + (tid * 4) % FT_N;
            const float4 v = *reinterpret_cast<const float4*>(sg_smem + tid * 4);
            const float vv[4] = {v.x, v.y, v.z, v.w};
            if (b < p.B) {
              if (fwd && l < L - 1) {
                // hidden layer: h_{l+1} = act(z), stored row-major and transposed
                const SgLayer& Nx = p.ly[l + 1];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const int j = c0 + c;
                  if (j < Y.out) {
                    const float a = act_fwd(vv[c], p.act);
                    Nx.h[(long long)b * Nx.ld + j] = a;
                    Nx.hT[(long long)j * p.Bp + b] = a;
                  }
                }
              } else if (fwd) {
                // head: residual, dU/d(out) and the per-output sum of squares (BNN_SGDMC.py:69-74, :83)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const int o = c0 + c;
                  if (o < Y.out) {
                    float dzv = 0.0f;
                    if (b < rows) {
                      const float r = vv[c] - __ldg(p.y + s_perm[b] * p.nout + o);
                      dzv = scale * expf(th[p.n_head + o]) * r;
                      atomicAdd(&p.acc[o], (double)r * (double)r);
                    }
                    Y.dz[(long long)b * Y.ldz + o] = dzv;
                    Y.dzT[(long long)o * p.Bp + b] = dzv;
                  }
                }
              } else {
                // dz_{l-1} = dh * act'(h_l)
                const SgLayer& Pv = p.ly[l - 1];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const int k = c0 + c;
                  if (k < Y.in) {
                    const float hv = __ldcg(Y.h + (long long)b * Y.ld + k);
                    const float d = vv[c] * act_bwd(hv, p.act);
                    Pv.dz[(long long)b * Pv.ldz + k] = d;
                    Pv.dzT[(long long)k * p.Bp + b] = d;
                  }
                }
              }
            }
          }
          __syncthreads();                 // the result tile in shared memory is free for the next job's loads
        } else {
          // ---- log-precision group (BNN_SGDMC.py:62, :96-98): closed-form gradient, same update rule, lr_noise
          if (tid < ((p.n_theta - p.n_head) >> 2)) {
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
                const float s2 = (float)__ldcg(&p.acc[o]);
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
        }
      }
      if (probe && tid == 0) probe[4] = clock64();
      if (p.trace && it == p.n_steps - 1 && tid == 0) p.trace[((long long)blockIdx.x * 2 * L + ph) * 2] = clock64();
      if (!grid_barrier(p.bar, bar_target, p.err, &s_flag)) return;
      if (p.trace && it == p.n_steps - 1 && tid == 0) p.trace[((long long)blockIdx.x * 2 * L + ph) * 2 + 1] = clock64();
    }

    // ================= end of step: loss of the last step, reset of the accumulators =================
    if (blockIdx.x == 0 && tid == 0) {
      if (want_loss) {
        float total_ll = 0.0f, total_lp = 0.0f;
        for (int o = 0; o < p.nout; ++o) {
          const float lp = th[p.n_head + o], prec = expf(lp);
          const float s2 = (float)__ldcg(&p.acc[o]);
          total_ll += -0.5f * prec * s2 + (float)rows * (0.5f * lp - 0.91893853320467274178f);
          total_lp += p.al * p.log_be + (p.al - 1.0f) * lp - p.be * prec - p.lgamma_al + lp;
        }
        const float quad = (float)__ldcg(&p.acc[p.nout]);
        const float logp = -0.5f * quad - p.log_const + total_lp;
        p.loss_out[0] = -(total_ll * scale + logp);
      }
      for (int o = 0; o <= p.nout; ++o) p.acc[o] = 0.0;
      __threadfence();
    }
    // the next accumulation into acc happens after L - 1 more grid barriers (or, for L == 1, after none: guard that case)
    if (L == 1 && !grid_barrier(p.bar, bar_target, p.err, &s_flag)) return;
  }
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
      Y.h[(long long)b * Y.ld + Y.in] = 1.0f;
      Y.hT[(long long)Y.in * p.Bp + b] = 1.0f;
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
  long long ws_bytes;
  int grid;
  size_t smem;
};

static int sg_plan(const BnnSgldChain* c, SgPlan* plan, bool bind) {
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
  int max_jobs = 1;
  for (int l = 0; l < p.L; ++l) {
    SgLayer& Y = p.ly[l];
    Y.in = Lo.in[l]; Y.out = Lo.out[l]; Y.ld = Lo.ld[l]; Y.off = Lo.off[l];
    Y.ldz = (int)r4(Y.out); Y.ldT = (int)r4(Y.out);
    Y.offT = oT;
    if (l > 0) oT += (long long)Y.in * Y.ldT;
    const double std = gain * sqrt(2.0 / (double)(Y.in + Y.out));
    Y.inv_var = (float)(1.0 / (std * std));
    log_const += (double)Y.in * Y.out * (log(std) + 0.5 * log(2.0 * 3.14159265358979323846));
    const int fwd = ((p.B + FT_M - 1) / FT_M) * ((Y.out + FT_N - 1) / FT_N);
    const int bwd = ((Y.out + WT_M - 1) / WT_M) * ((Y.ld + WT_N - 1) / WT_N) +
                    (l > 0 ? ((p.B + FT_M - 1) / FT_M) * ((Y.in + FT_N - 1) / FT_N) : 0) + 1;
    if (fwd > max_jobs) max_jobs = fwd;
    if ((bwd + 1) / 2 > max_jobs) max_jobs = (bwd + 1) / 2;   // backward phases may take two rounds
  }
  plan->nT = oT > 0 ? oT : 4;
  // workspace: per layer h, hT (l >= 1), dz, dzT
  for (int l = 0; l < p.L; ++l) {
    SgLayer& Y = p.ly[l];
    if (l > 0) { ow += r4((long long)p.Bp * Y.ld); ow += r4((long long)Y.ld * p.Bp); }
    ow += r4((long long)p.Bp * Y.ldz);
    ow += r4((long long)Y.out * p.Bp);
  }
  plan->ws_floats = ow;
  const long long tail = 8LL * (p.nout + 1) + 64;          // acc doubles, barrier counter, error flag
  plan->ws_bytes = ((ow * 4 + 15) & ~15LL) + tail;
  p.log_const = (float)log_const;
  p.lr_w = c->lr_weight; p.lr_n = c->lr_noise; p.alpha = c->precond_decay; p.lam = c->diagonal_bias;
  p.al = c->alpha_n; p.be = c->beta_n; p.lgamma_al = lgammaf(c->alpha_n); p.log_be = logf(c->beta_n);
  p.precond_mode = c->precond_mode;
  p.seed = c->seed;
  plan->smem = (size_t)kSgStagesW * (WT_M + WT_N) * kSgLds * 4;
  if ((size_t)kSgStagesF * (FT_M + FT_N) * kSgLds * 4 > plan->smem) plan->smem = (size_t)kSgStagesF * (FT_M + FT_N) * kSgLds * 4;
  int grid = sm_count();
  if (max_jobs < grid) grid = max_jobs;
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
      if (l > 0) { Y.h = w + o; o += r4((long long)p.Bp * Y.ld); Y.hT = w + o; o += r4((long long)Y.ld * p.Bp); }
      Y.dz = w + o; o += r4((long long)p.Bp * Y.ldz);
      Y.dzT = w + o; o += r4((long long)Y.out * p.Bp);
    }
    char* tailp = reinterpret_cast<char*>(c->workspace) + ((ow * 4 + 15) & ~15LL);
    p.acc = reinterpret_cast<double*>(tailp);
    p.bar = reinterpret_cast<unsigned int*>(tailp + 8LL * (p.nout + 1));
    p.err = reinterpret_cast<int*>(tailp + 8LL * (p.nout + 1) + 16);
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

static long long* g_sg_trace = nullptr;   // debug: set by bnn_sgld_trace
static int g_sg_trace_grid = 0, g_sg_trace_L = 0;

extern "C" int32_t bnn_sgld_run(const BnnSgldChain* c, int32_t cur, int64_t t0, int64_t perm_off, int32_t n_steps, float* loss_out,
                                float* grad_out, void* stream) {
  if (!c || (cur != 0 && cur != 1) || n_steps < 1 || t0 < 0 || perm_off < 0) BNN_FAIL("bnn_sgld_run: bad arguments");
  if (!c->X || !c->y || !c->perm || c->num_train < 1) BNN_FAIL("bnn_sgld_run: X, y, perm and num_train are required");
  if (grad_out && n_steps != 1) BNN_FAIL("bnn_sgld_run: the gradient-only mode runs exactly one step");
  if (perm_off + (int64_t)(n_steps - 1) * c->batch >= c->num_train)
    BNN_FAIL("bnn_sgld_run: %d steps of batch %d from offset %lld run past the epoch (%lld rows)", n_steps, c->batch,
             (long long)perm_off, (long long)c->num_train);
  SgPlan plan;
  if (sg_plan(c, &plan, true)) return -1;
  SgParams& p = plan.p;
  p.X = c->X; p.y = c->y; p.perm = (const long long*)c->perm; p.num_train = c->num_train; p.perm_off = perm_off;
  p.t0 = t0; p.n_steps = n_steps; p.cur = cur;
  p.loss_out = loss_out; p.grad_out = grad_out;
  cudaStream_t st = (cudaStream_t)stream;
  // accumulators, barrier counter and error flag start at zero for every launch
  BNN_CUDA(cudaMemsetAsync(p.acc, 0, (size_t)(8LL * (p.nout + 1) + 64), st));
  BNN_CUDA(cudaFuncSetAttribute(sgld_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  int per_sm = 0;
  BNN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sgld_step_kernel, kSgThreads, plan.smem));
  if (per_sm < 1) BNN_FAIL("bnn_sgld_run: the step kernel does not fit on an SM");
  int grid = plan.grid;
  if (grid > per_sm * sm_count()) grid = per_sm * sm_count();
  p.trace = g_sg_trace;
  if (const char* e = getenv("BNN_SGLD_DBG")) p.dbg = atoi(e);
  g_sg_trace_grid = grid; g_sg_trace_L = p.L;
  void* args[] = {(void*)&p};
  BNN_CUDA(cudaLaunchCooperativeKernel((void*)sgld_step_kernel, dim3((unsigned)grid), dim3(kSgThreads), args, plan.smem, st));
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

// debug: enable = 1 allocates the per-CTA phase timeline written by the following bnn_sgld_run launches (last step of each
// launch); with host_out != NULL copies it out: [grid][2 L][2] clock64 values (work done, barrier passed); returns grid in *grid_out
extern "C" int32_t bnn_sgld_trace(int32_t enable, long long* host_out, int32_t* grid_out, int32_t* phases_out) {
  if (enable && !g_sg_trace) {
    BNN_CUDA(cudaMalloc(&g_sg_trace, (size_t)(1024 * 2 * BNN_MAX_LINEAR * 2 + 256) * sizeof(long long)));
    BNN_CUDA(cudaMemset(g_sg_trace, 0, (size_t)(1024 * 2 * BNN_MAX_LINEAR * 2 + 256) * sizeof(long long)));
  }
  if (host_out && g_sg_trace) {
    BNN_CUDA(cudaDeviceSynchronize());
    BNN_CUDA(cudaMemcpy(host_out, g_sg_trace, (size_t)g_sg_trace_grid * 2 * g_sg_trace_L * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    // fine-grained probes of CTA 1 (first job of each phase): appended behind the coarse timeline
    BNN_CUDA(cudaMemcpy(host_out + (size_t)g_sg_trace_grid * 2 * g_sg_trace_L * 2, g_sg_trace + 1024 * 2 * BNN_MAX_LINEAR * 2,
                        (size_t)2 * g_sg_trace_L * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  if (grid_out) *grid_out = g_sg_trace_grid;
  if (phases_out) *phases_out = 2 * g_sg_trace_L;
  if (!enable && g_sg_trace) { cudaFree(g_sg_trace); g_sg_trace = nullptr; }
  return 0;
}

// 0 = ok; 1 = a grid barrier of the last launches timed out (never expected; the chain state is then undefined)
extern "C" int32_t bnn_sgld_status(const BnnSgldChain* c, int32_t* status_host, void* stream) {
  if (!c || !status_host) BNN_FAIL("bnn_sgld_status: bad arguments");
  SgPlan plan;
  if (sg_plan(c, &plan, true)) return -1;
  BNN_CUDA(cudaMemcpyAsync(status_host, plan.p.err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  BNN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
