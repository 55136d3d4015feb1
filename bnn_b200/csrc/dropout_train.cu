// Row f2 of SURVEY.md 8, second half: the MC-dropout minibatch training step of BNN_Dropout.train (BNN_Dropout.py:39-65) for
// networks that fit ONE SM, as a persistent single-CTA kernel that runs n_steps steps per launch:
//   minibatch gather (:53-56) -> NN_Dropout.forward (:15-21: every Linear whose input is wider than one unit sees its input
//   multiplied by a fresh Bernoulli(1 - p) keep mask; F.dropout(x, p) * (1 - p)) -> MSELoss(reduction='sum') (:55,59) ->
//   backward -> Adam with weight decay l2_reg on the weights, none on the biases (:43-52,61).
// Same structure as csrc/bbb_train.cu: the packed network (bias-last rows, the layout of the forward path), its two Adam moment
// vectors, the masked activations and two gradient panels live in shared memory for the whole launch.  The keep masks are never
// stored: forward and backward evaluate the same counter-based draw (Philox stream 6, subsequence = step) or read the injected
// mask (tests: the reference's own F.dropout draws).  The parameter gradient is consumed by Adam in the thread that forms it.
// The reference's cfg 3 network (9-100-100-1, batch 32: 11,704 packed parameters) needs 207 KB (rows at the conflict-free pitch).
//
// The same kernel, instantiated with CONCRETE = true, is the concrete-dropout training step of BNN_CDropout.train
// (BNN_CDropout.py:114-139): CDropout.forward (:21-34) scales the input of EVERY Linear by the relaxed keep factor
//   keep = 1 - sigmoid((log(p + e) - log(1 - p + e) + log(u + e) - log(1 - u + e)) / 0.1),  e = 1e-7, u ~ U[0, 1),
//   p = clamp_probs(sigmoid(p_logit)) (:18-19), one learnable p_logit per layer;
// loss = MSELoss(sum) + (wreg - ent) * rows / num_train with (BNN_CDropout.reg, :141-151; CDropoutLinear.reg, :53-57)
//   wreg = sum_l lscale^2 (1 - p_l) noise_level^2 sum(W_l^2),  ent = sum_l 2 dr noise_level^2 in_l H(p_l);
// Adam without weight decay on W, b and p_logit (:116).  noise_level is a device scalar: the launch that ends an epoch sets it to
// max(min_noise, sqrt(sum of the launch's squared errors / num_train)) (:134) -- no host round trip per epoch.
// Gradients in closed form: d keep / d z = -(1 - keep) keep / T, so d L / d p_logit_l = sum_{b,k} -g[b,k] xm[b,k] (1 - keep) / T
// (xm = the stored masked input, g = the gradient with respect to it) times dz/dp dp/dlogit, plus the regulariser's
// -(rows / N) noise_level^2 (lscale^2 sum(W^2) + 2 dr in log((1 - p) / p)) dp/dlogit.
#include <math.h>

#include "common.cuh"
#include "philox.cuh"

namespace bnn {

#ifndef BNN_DT_THREADS
#define BNN_DT_THREADS 1024
#endif
constexpr int kDtThreads = BNN_DT_THREADS;
constexpr uint32_t STREAM_DROPOUT_TRAIN = 6;

struct DtLayer {
  int in, out, ld, masked;
  int lp;             // shared-memory row pitch of the layer's weight rows AND of its input rows: ld rounded up so that lp / 4 is
                      // odd -- eight threads reading float4s of eight different rows then hit eight different bank groups
  int half;           // ceil(out / 2): a thread of the forward / parameter phases owns output units j and j + half
  int k4p;            // float4 columns of the input, rounded up to a multiple of 8 (lane mapping of the input-gradient phase)
  long long off;      // float offset of the layer in the packed vector (global memory, row pitch ld)
  int soff;           // float offset of the layer in the shared-memory copies of W / m / v (row pitch lp)
  int a_off;          // shared-memory float offset of this layer's (masked, augmented) input [Bp][lp]
  long long m_off;    // offset of this layer's keep mask in the per-step mask block ([B][in]); -1 = not masked
};

struct DtParams {
  int L, act, B, nout, D;
  int Bp, gp;                               // batch rounded up to 4; pitch of the transposed gradient panels (Bp + 4: gp / 4 odd)
  int sP;                                   // floats of one shared-memory parameter vector (pitch lp)
  DtLayer ly[BNN_MAX_LINEAR];
  long long stride, mask_per_step;
  float* W; float* adam;                    // global: [stride], [2][stride] = m, v
  const float* X; const float* y; const long long* perm;
  long long num_train, perm_off, t0;
  int n_steps;
  float lr, beta1, beta2, eps, l2_reg, p_drop;
  unsigned long long seed;
  float* loss_out;                          // [n_steps] or null: sum of squared errors of the step's batch, before its update
  const float* mask_in;                     // tests: injected keep factors (concrete: uniforms) [n_steps][mask_per_step]; null = Philox
  int sm_act, sm_g0, sm_g1, sm_y, sm_red, sm_cd, max_w;
  // concrete dropout only
  float* plogit; float* padam;              // global: [L], [2][L]
  float* noise_level;                       // device scalar, in / out
  float* stats_out;                         // [n_steps][3] or null: squared errors, wreg * rows / N, ent * rows / N
  float lscale2, dr, min_noise;
  int epoch_end;
};

constexpr float kCdEps = 1e-7f, kCdInvTemp = 10.0f;                  // BNN_CDropout.py:23-24
constexpr float kProbEps = 1.1920928955078125e-07f;                   // torch clamp_probs: finfo(float32).eps

// per-layer concrete-dropout state in shared memory (floats): [0] p_logit [1] m [2] v [3] p [4] log(p+e) - log(1-p+e) [5] grad
constexpr int kCdStride = 8;

// keep factors of input units k0 .. k0 + 3 (k0 a multiple of 4) of row b: ONE Philox block when the row's draws start on a block
// boundary (in % 4 == 0: every hidden width that matters), two otherwise -- a draw per element would use one word of every block
template <bool CONCRETE>
__device__ __forceinline__ void dt_keep4(const DtParams& p, const DtLayer& Y, int it, int b, int k0, float lpr, float keep[4]) {
  const long long u0 = Y.m_off + (long long)b * Y.in + k0;
  float r[4];
  if (p.mask_in) {
    const float* src = p.mask_in + (long long)it * p.mask_per_step + u0;
#pragma unroll
    for (int c = 0; c < 4; ++c) r[c] = k0 + c < Y.in ? src[c] : 0.0f;
  } else {
    const uint32_t sub = (uint32_t)(p.t0 + it);
    const u32x4 w0 = philox_block(p.seed, sub, STREAM_DROPOUT_TRAIN, (uint64_t)(u0 >> 2));
    const int sh = (int)(u0 & 3);
    if (sh == 0) {
      r[0] = u24(w0.x); r[1] = u24(w0.y); r[2] = u24(w0.z); r[3] = u24(w0.w);
    } else {
      const u32x4 w1 = philox_block(p.seed, sub, STREAM_DROPOUT_TRAIN, (uint64_t)(u0 >> 2) + 1);
#pragma unroll
      for (int c = 0; c < 4; ++c) r[c] = u24(sh + c < 4 ? pick_word(w0, sh + c) : pick_word(w1, sh + c - 4));
    }
  }
  if constexpr (CONCRETE) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float z = (lpr + logf(r[c] + kCdEps) - logf(1.0f - r[c] + kCdEps)) * kCdInvTemp;
      keep[c] = 1.0f - 1.0f / (1.0f + expf(-z));
    }
  } else {
    // injected masks are the keep factors themselves; drawn ones are Bernoulli(1 - p) of the 24-bit uniforms
#pragma unroll
    for (int c = 0; c < 4; ++c) keep[c] = p.mask_in ? r[c] : bernoulli_keep(r[c], p.p_drop);
  }
}

__device__ __forceinline__ float dt_act_bwd(float a, int act) {   // derivative through the activation's OUTPUT
  switch (act) {
    case BNN_ACT_RELU: return a > 0.0f ? 1.0f : 0.0f;
    case BNN_ACT_TANH: return 1.0f - a * a;
    case BNN_ACT_SIGMOID: return a * (1.0f - a);
    default: return 1.0f;
  }
}

// ---- the kernel ------------------------------------------------------------------------------------------------------------
// Thread tiles (the shared-memory pipe, not the FMA pipe, bounds a one-SM step: every phase is arranged so that a thread issues
// at most ~6 LDS.128 per 32 FMAs, and so that the float4s a quarter-warp reads are either one address or eight bank groups):
//   forward    out[b][j]:  4 rows b x 2 units (j, j + half); consecutive threads = consecutive j  (x: broadcast, W: 8 rows)
//   input grad dh[b][k]:   4 rows b x 4 columns k, the sum over j split over 4 lanes (lane bits 3-4) and folded by shuffles;
//                          lane bits 0-2 = consecutive float4 columns (W: contiguous, gT: broadcast)
//   parameters dW[j][k]:   2 units (j, j + half) x 4 columns k over the batch in steps of 4 rows, then Adam on those 8 entries
// The gradient panels are kept TRANSPOSED (gT[unit][row], pitch gp) so that both backward phases read them as float4s.
template <bool CONCRETE>
__global__ void __launch_bounds__(kDtThreads, 1) dropout_train_kernel(const __grid_constant__ DtParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31;
  constexpr int kWarps = kDtThreads / 32;
  const int sP = p.sP;
  float* s_w = sm;
  float* s_m = s_w + sP;
  float* s_v = s_m + sP;
  float* s_act = sm + p.sm_act;
  float* s_ga = sm + p.sm_g0;   // dL/d(out) of the current layer, transposed [max_w][gp] ...
  float* s_gb = sm + p.sm_g1;   // ... and of the layer below (ping-pong)
  float* s_y = sm + p.sm_y;     // [B][nout] targets, then [B][nout] predictions
  double* s_red = reinterpret_cast<double*>(sm + p.sm_red);   // [warps]: loss partial sums
  double* s_redp = s_red + kWarps;                            // [warps]: p_logit gradient partial sums (concrete)
  double* s_redw = s_redp + kWarps;                           // [L][warps]: sum(W^2) partial sums (concrete)
  double* s_w2 = s_redw + BNN_MAX_LINEAR * kWarps;            // [L] sum(W_l^2) at the start of the step; [L] = the launch's loss sum
  float* s_cd = sm + p.sm_cd;                                 // [L][kCdStride]
  const int L = p.L, B = p.B, Bp = p.Bp, gp = p.gp;
  float* s_pred = s_y + B * p.nout;

  // parameters and moments: global rows of pitch ld -> shared rows of pitch lp (pad columns zero)
  for (int i = tid; i < 3 * sP; i += kDtThreads) sm[i] = 0.0f;
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const DtLayer& Y = p.ly[l];
    for (int i = tid; i < Y.out * Y.ld; i += kDtThreads) {
      const int j = i / Y.ld, k = i - j * Y.ld;
      const long long g = Y.off + i;
      const int d = Y.soff + j * Y.lp + k;
      s_w[d] = p.W[g]; s_m[d] = p.adam[g]; s_v[d] = p.adam[p.stride + g];
    }
  }
  float nv = 1.0f;
  if (CONCRETE) {
    if (tid < L) { s_cd[tid * kCdStride] = p.plogit[tid]; s_cd[tid * kCdStride + 1] = p.padam[tid]; s_cd[tid * kCdStride + 2] = p.padam[L + tid]; }
    if (tid == 0) s_w2[L] = 0.0;
    const float nl = *p.noise_level;
    nv = nl * nl;
  }
  __syncthreads();
  double b1t = pow((double)p.beta1, (double)p.t0), b2t = pow((double)p.beta2, (double)p.t0);

#pragma unroll 1
  for (int it = 0; it < p.n_steps; ++it) {
    const long long boff = p.perm_off + (long long)it * B;
    const long long left = p.num_train - boff;
    const int rows = left < B ? (int)left : B;                     // the last batch of an epoch may be short
    const float frac = (float)((double)rows / (double)p.num_train);
    if (CONCRETE) {
      // the step's dropout probabilities (BNN_CDropout.py:18-19) and sum(W_l^2) of the weights before the update
      if (tid < L) {
        float* c = s_cd + tid * kCdStride;
        const float sg = 1.0f / (1.0f + expf(-c[0]));
        const float pr = fminf(fmaxf(sg, kProbEps), 1.0f - kProbEps);
        c[3] = pr;
        c[4] = logf(pr + kCdEps) - logf(1.0f - pr + kCdEps);
      }
      for (int l = 0; l < L; ++l) {
        const DtLayer& Y = p.ly[l];
        double part = 0.0;
        for (int i = tid; i < Y.out * Y.lp; i += kDtThreads) {
          const int k = i % Y.lp;
          const float w = s_w[Y.soff + i];
          if (k < Y.in) part += (double)(w * w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s_redw[l * kWarps + (tid >> 5)] = part;
      }
      __syncthreads();
      if (tid < L) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += s_redw[tid * kWarps + w];
        s_w2[tid] = s;
      }
    }
    // ---- gather: masked, augmented rows (x * keep | 1 | 0) of the batch, and the targets
    {
      const DtLayer& Y0 = p.ly[0];
      float* a0 = s_act + Y0.a_off;
      const float lpr = CONCRETE ? s_cd[4] : 0.0f;
      const int q0 = Y0.lp >> 2;
      for (int i = tid; i < Bp * q0; i += kDtThreads) {
        const int b = i / q0, k0 = (i - b * q0) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (b < rows) {
          const float* xr = p.X + __ldg(p.perm + boff + b) * p.D;
#pragma unroll
          for (int c = 0; c < 4; ++c) v[c] = k0 + c < p.D ? __ldg(xr + k0 + c) : (k0 + c == p.D ? 1.0f : 0.0f);
          if (Y0.masked && k0 < p.D) {
            float keep[4];
            dt_keep4<CONCRETE>(p, Y0, it, b, k0, lpr, keep);
#pragma unroll
            for (int c = 0; c < 4; ++c) if (k0 + c < p.D) v[c] *= keep[c];
          }
        }
        *reinterpret_cast<float4*>(a0 + b * Y0.lp + k0) = make_float4(v[0], v[1], v[2], v[3]);
      }
      for (int i = tid; i < B * p.nout; i += kDtThreads) {
        const int b = i / p.nout;
        s_y[i] = b < rows ? __ldg(p.y + __ldg(p.perm + boff + b) * p.nout + (i - b * p.nout)) : 0.0f;
      }
    }
    __syncthreads();
    // ---- forward
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const DtLayer& Y = p.ly[l];
      const float* a = s_act + Y.a_off;
      const bool last = l + 1 == L;
      const float lpr = (CONCRETE && !last) ? s_cd[(l + 1) * kCdStride + 4] : 0.0f;
      const int n_item = (Bp >> 2) * Y.half;
      for (int i = tid; i < n_item; i += kDtThreads) {
        const int bq = i / Y.half, j0 = i - bq * Y.half, j1 = j0 + Y.half;
        const bool two = j1 < Y.out;
        const float* ar = a + (bq * 4) * Y.lp;
        const float* w0 = s_w + Y.soff + j0 * Y.lp;
        const float* w1 = two ? w0 + Y.half * Y.lp : w0;
        float o[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) { o[r][0] = 0.0f; o[r][1] = 0.0f; }
        for (int k = 0; k < Y.lp; k += 4) {
          const float4 wa = *reinterpret_cast<const float4*>(w0 + k);
          const float4 wb = *reinterpret_cast<const float4*>(w1 + k);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float4 x = *reinterpret_cast<const float4*>(ar + r * Y.lp + k);
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
            const int j = c ? j1 : j0;
            if (c && !two) continue;
            float v = o[r][c];
            if (last) {
              if (b < B) s_pred[b * p.nout + j] = b < rows ? v : 0.0f;
            } else {
              const DtLayer& N = p.ly[l + 1];
              s_act[N.a_off + b * N.lp + j] = b < rows ? apply_act(v, p.act) : 0.0f;      // the keep factors follow below
            }
          }
        }
      }
      __syncthreads();
      if (!last) {          // the next layer's input: keep factors on its units (one Philox block per four), ones column, zero pad
        const DtLayer& N = p.ly[l + 1];
        float* an = s_act + N.a_off;
        const int qn = N.lp >> 2;
        for (int i = tid; i < Bp * qn; i += kDtThreads) {
          const int b = i / qn, k0 = (i - b * qn) * 4;
          const float4 x = *reinterpret_cast<const float4*>(an + b * N.lp + k0);
          float v[4] = {x.x, x.y, x.z, x.w};
          float keep[4] = {1.f, 1.f, 1.f, 1.f};
          if (N.masked && b < rows && k0 < N.in) dt_keep4<CONCRETE>(p, N, it, b, k0, lpr, keep);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int k = k0 + c;
            v[c] = b < rows ? (k < N.in ? v[c] * keep[c] : (k == N.in ? 1.0f : 0.0f)) : 0.0f;
          }
          *reinterpret_cast<float4*>(an + b * N.lp + k0) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
      }
    }
    // ---- loss = sum (pred - y)^2 (MSELoss(reduction='sum'), BNN_Dropout.py:55); dL/d(pred) -> s_ga (transposed)
    {
      double part = 0.0;
      for (int i = tid; i < Bp * p.nout; i += kDtThreads) {
        const int b = i / p.nout, o = i - b * p.nout;
        const float r = b < rows ? s_pred[i] - s_y[i] : 0.0f;
        part += (double)r * (double)r;
        s_ga[o * gp + b] = 2.0f * r;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) s_red[tid >> 5] = part;
    }
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < kWarps; ++w) s += s_red[w];
      if (p.loss_out) p.loss_out[it] = (float)s;
      if (CONCRETE) {
        s_w2[L] += (double)(float)s;
        if (p.stats_out) {                                           // BNN_CDropout.py:127-129,141-151
          float wreg = 0.0f, ent = 0.0f;
          for (int l = 0; l < L; ++l) {
            const float pr = s_cd[l * kCdStride + 3];
            wreg += p.lscale2 * (1.0f - pr) * nv * (float)s_w2[l];
            ent += p.dr * 2.0f * nv * ((float)p.ly[l].in * -(pr * logf(pr) + (1.0f - pr) * logf(1.0f - pr)));
          }
          p.stats_out[3 * it] = (float)s; p.stats_out[3 * it + 1] = wreg * frac; p.stats_out[3 * it + 2] = ent * frac;
        }
      }
    }
    // ---- backward + Adam, layer by layer from the top
    b1t *= (double)p.beta1; b2t *= (double)p.beta2;
    const float step_size = (float)((double)p.lr / (1.0 - b1t));
    const float bc2_sqrt = (float)sqrt(1.0 - b2t);
    const float om1 = 1.0f - p.beta1, om2 = 1.0f - p.beta2;
    float* g0 = s_ga;
    float* g1 = s_gb;
#pragma unroll 1
    for (int l = L - 1; l >= 0; --l) {
      const DtLayer& Y = p.ly[l];
      const float* a = s_act + Y.a_off;
      // gradient of the layer's (unmasked, pre-dropout) input, through the keep mask and the activation below: needs the
      // weights BEFORE their update.  Concrete dropout also needs it at layer 0: the masked input's gradient drives p_logit.
      if (l > 0 || CONCRETE) {
        const float lpr = CONCRETE ? s_cd[l * kCdStride + 4] : 0.0f;
        const int k4n = (Y.in + 3) >> 2, k8 = Y.k4p >> 3;
        const int n_item = 32 * k8 * (Bp >> 2);                      // lane = k4 & 7 | part << 3; above: k4 >> 3, then the row quad
        float gpart = 0.0f;
        for (int i = tid; i < n_item; i += kDtThreads) {
          const int part = (i >> 3) & 3, r = i >> 5;
          const int k4 = (i & 7) + 8 * (r % k8), bq = r / k8;
          const bool live = k4 < k4n;
          float g[4][4];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int c = 0; c < 4; ++c) g[rr][c] = 0.0f;
          if (live) {
            const float* wp = s_w + Y.soff + k4 * 4;
            const float* gq = g0 + bq * 4;
            for (int j = part; j < Y.out; j += 4) {
              const float4 w = *reinterpret_cast<const float4*>(wp + j * Y.lp);
              const float4 gj = *reinterpret_cast<const float4*>(gq + j * gp);
              const float gs[4] = {gj.x, gj.y, gj.z, gj.w};
#pragma unroll
              for (int rr = 0; rr < 4; ++rr) {
                g[rr][0] = fmaf(gs[rr], w.x, g[rr][0]); g[rr][1] = fmaf(gs[rr], w.y, g[rr][1]);
                g[rr][2] = fmaf(gs[rr], w.z, g[rr][2]); g[rr][3] = fmaf(gs[rr], w.w, g[rr][3]);
              }
            }
          }
          // fold the four partial sums (lanes differing in bits 3 and 4); afterwards lane `part` finishes ROW bq * 4 + part over the
          // tile's four columns: their keep factors are one Philox block
          float mine[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float v = g[rr][c];
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              if (rr == part) mine[c] = v;
            }
          }
          if (live) {
            const int b = bq * 4 + part, k0 = k4 * 4;
            const float4 x4 = *reinterpret_cast<const float4*>(a + b * Y.lp + k0);                // act(z) * keep
            const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
            float keep[4] = {1.f, 1.f, 1.f, 1.f};
            if (Y.masked) {
              if (b < rows) dt_keep4<CONCRETE>(p, Y, it, b, k0, lpr, keep);
              else keep[0] = keep[1] = keep[2] = keep[3] = 0.0f;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (k0 + c < Y.in) {
                if (CONCRETE) gpart -= mine[c] * xs[c] * (1.0f - keep[c]);          // d keep / d z = -(1 - keep) keep / T; xs = act * keep
                // where keep is not 0 the unmasked activation is xs / keep (Bernoulli: keep is 0 or within an ulp of 1;
                // concrete: keep = 1 - sigmoid is 0 or at least 2^-24)
                if (l > 0) g1[(k0 + c) * gp + b] = keep[c] != 0.0f ? mine[c] * keep[c] * dt_act_bwd(xs[c] / keep[c], p.act) : 0.0f;
              }
            }
          }
        }
        if (CONCRETE) {
          double part = (double)gpart;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          if (lane == 0) s_redp[tid >> 5] = part;
        }
      }
      __syncthreads();
      if (CONCRETE && tid == 0) {
        float* c = s_cd + l * kCdStride;
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += s_redp[w];
        const float pr = c[3];
        const float sg = 1.0f / (1.0f + expf(-c[0]));
        const float dpdl = (sg >= kProbEps && sg <= 1.0f - kProbEps) ? sg * (1.0f - sg) : 0.0f;   // clamp_probs passes no gradient outside
        const float dzdp = 1.0f / (pr + kCdEps) + 1.0f / (1.0f - pr + kCdEps);
        const float greg = frac * nv * (-p.lscale2 * (float)s_w2[l] - 2.0f * p.dr * (float)Y.in * (logf(1.0f - pr) - logf(pr)));
        c[5] = ((float)s * kCdInvTemp * dzdp + greg) * dpdl;
      }
      // parameters: thread = units (j, j + half) x four columns: gradient over the batch, weight decay, Adam, in place
      {
        const int q4 = Y.lp >> 2;
        const float decay = CONCRETE ? 2.0f * frac * p.lscale2 * (1.0f - s_cd[l * kCdStride + 3]) * nv : p.l2_reg;
        const int n_item = Y.half * q4;
        for (int i = tid; i < n_item; i += kDtThreads) {
          const int j0 = i / q4, k = (i - j0 * q4) * 4, j1 = j0 + Y.half;
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
              const float4 x = *reinterpret_cast<const float4*>(a + (b + rr) * Y.lp + k);
              d[0][0] = fmaf(us[rr], x.x, d[0][0]); d[0][1] = fmaf(us[rr], x.y, d[0][1]);
              d[0][2] = fmaf(us[rr], x.z, d[0][2]); d[0][3] = fmaf(us[rr], x.w, d[0][3]);
              d[1][0] = fmaf(vs[rr], x.x, d[1][0]); d[1][1] = fmaf(vs[rr], x.y, d[1][1]);
              d[1][2] = fmaf(vs[rr], x.z, d[1][2]); d[1][3] = fmaf(vs[rr], x.w, d[1][3]);
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h && !two) continue;
            const int w0 = Y.soff + (h ? j1 : j0) * Y.lp + k;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (k + c <= Y.in) {
                const int w = w0 + c;
                const float wv = s_w[w];
                // Adam's weight_decay (BNN_Dropout.py:43-51) / the gradient of wreg (BNN_CDropout.py:148): weights only
                const float gr = k + c < Y.in ? fmaf(decay, wv, d[h][c]) : d[h][c];
                const float m = p.beta1 * s_m[w] + om1 * gr, v = p.beta2 * s_v[w] + om2 * gr * gr;
                s_m[w] = m; s_v[w] = v;
                s_w[w] = wv - step_size * (m / (sqrtf(v) / bc2_sqrt + p.eps));
              }
            }
          }
        }
      }
      __syncthreads();
      float* t = g0; g0 = g1; g1 = t;
    }
    if (CONCRETE && tid < L) {                                        // Adam on the p_logits; read again by this thread next step
      float* c = s_cd + tid * kCdStride;
      const float gr = c[5];
      const float m = p.beta1 * c[1] + om1 * gr, v = p.beta2 * c[2] + om2 * gr * gr;
      c[1] = m; c[2] = v;
      c[0] -= step_size * (m / (sqrtf(v) / bc2_sqrt + p.eps));
    }
  }
  for (int l = 0; l < L; ++l) {
    const DtLayer& Y = p.ly[l];
    for (int i = tid; i < Y.out * Y.ld; i += kDtThreads) {
      const int j = i / Y.ld, k = i - j * Y.ld;
      const long long g = Y.off + i;
      const int d = Y.soff + j * Y.lp + k;
      p.W[g] = s_w[d]; p.adam[g] = s_m[d]; p.adam[p.stride + g] = s_v[d];
    }
  }
  if (CONCRETE) {
    if (tid < L) { p.plogit[tid] = s_cd[tid * kCdStride]; p.padam[tid] = s_cd[tid * kCdStride + 1]; p.padam[L + tid] = s_cd[tid * kCdStride + 2]; }
    if (tid == 0 && p.epoch_end)                                      // BNN_CDropout.py:134
      *p.noise_level = fmaxf(p.min_noise, (float)sqrt(s_w2[L] / (double)p.num_train));
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
struct DtPlan {
  DtParams p;
  size_t smem;
};

// 0 = ok, 1 = does not fit one SM (message set), < 0 = bad arguments
static int dt_plan(const BnnMlpDesc* d, int batch, bool concrete, DtPlan* plan) {
  Layout Lo;
  if (make_layout(d, &Lo)) return -1;
  if (batch < 1 || batch > 1024) { set_error("bnn_dropout_train: batch must be in [1, 1024] (got %d)", batch); return -1; }
  DtParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.L = Lo.n_lin; p.act = Lo.act; p.B = batch; p.nout = Lo.out[Lo.n_lin - 1]; p.D = Lo.in[0];
  p.stride = (Lo.off[Lo.n_lin] + 3) & ~3LL;
  p.Bp = (batch + 3) & ~3;
  p.gp = p.Bp + 4 - (p.Bp & 4);                                    // gp / 4 odd
  if (((p.gp >> 2) & 1) == 0) p.gp += 4;
  long long act = 0, m = 0, sP = 0;
  int max_w = p.nout;
  for (int l = 0; l < p.L; ++l) {
    DtLayer& Y = p.ly[l];
    Y.in = Lo.in[l]; Y.out = Lo.out[l]; Y.ld = Lo.ld[l]; Y.off = Lo.off[l];
    Y.lp = ((Y.ld >> 2) & 1) ? Y.ld : Y.ld + 4;
    Y.half = (Y.out + 1) >> 1;
    Y.k4p = (((Y.in + 3) >> 2) + 7) & ~7;
    Y.soff = (int)sP; sP += (long long)Y.out * Y.lp;
    Y.masked = (concrete || Y.in > 1) ? 1 : 0;                     // BNN_Dropout.py:18: "x.shape[1] > 1"; CDropout: every layer
    Y.a_off = (int)act; act += (long long)p.Bp * Y.lp;
    Y.m_off = Y.masked ? m : -1;
    if (Y.masked) m += ((long long)batch * Y.in + 3) & ~3LL;
    if (Y.out > max_w) max_w = Y.out;
    if (Y.in + 1 > max_w) max_w = Y.in + 1;
  }
  p.sP = (int)sP;
  p.mask_per_step = m;
  p.max_w = max_w;
  long long f = 3 * sP;
  p.sm_act = (int)f; f += (act + 3) & ~3LL;
  p.sm_g0 = (int)f; f += (long long)max_w * p.gp;
  p.sm_g1 = (int)f; f += (long long)max_w * p.gp;
  p.sm_y = (int)f; f += (2LL * p.Bp * p.nout + 3) & ~3LL;
  p.sm_red = (int)f; f += 2 * ((2 + BNN_MAX_LINEAR) * (kDtThreads / 32) + BNN_MAX_LINEAR + 1);
  p.sm_cd = (int)f; f += BNN_MAX_LINEAR * kCdStride;
  plan->smem = (size_t)f * 4;
  if (plan->smem > (size_t)227 * 1024) {
    set_error("bnn_(c)dropout_train: %lld packed parameters at batch %d need %zu bytes of shared memory (one SM has 227 KB)",
              (long long)p.stride, batch, plan->smem);
    return 1;
  }
  return 0;
}

}  // namespace bnn

using namespace bnn;

extern "C" int32_t bnn_dropout_train_supported(const BnnMlpDesc* d, int32_t batch) {
  if (!d) return 0;
  DtPlan plan;
  return dt_plan(d, batch, false, &plan) == 0;
}

extern "C" int64_t bnn_dropout_train_mask_per_step(const BnnMlpDesc* d, int32_t batch) {
  if (!d) { set_error("null desc"); return -1; }
  DtPlan plan;
  if (dt_plan(d, batch, false, &plan) < 0) return -1;
  return plan.p.mask_per_step;
}

extern "C" int32_t bnn_dropout_train_run(const BnnDropoutTrain* c, int64_t t0, int64_t perm_off, int32_t n_steps, float* loss_out,
                                         const float* mask_in, void* stream) {
  if (!c || n_steps < 1 || t0 < 0 || perm_off < 0) BNN_FAIL("bnn_dropout_train_run: bad arguments");
  if (!c->W || !c->adam || !c->X || !c->y || !c->perm || c->num_train < 1)
    BNN_FAIL("bnn_dropout_train_run: W, adam, X, y, perm and num_train are required");
  if (!(c->lr > 0.f) || c->p_drop < 0.f || c->p_drop >= 1.f) BNN_FAIL("bnn_dropout_train_run: lr must be positive and p_drop in [0, 1)");
  if (perm_off + (int64_t)(n_steps - 1) * c->batch >= c->num_train)
    BNN_FAIL("bnn_dropout_train_run: %d steps of batch %d from offset %lld run past the epoch (%lld rows)", n_steps, c->batch,
             (long long)perm_off, (long long)c->num_train);
  DtPlan plan;
  if (dt_plan(&c->desc, c->batch, false, &plan)) return -1;
  DtParams& p = plan.p;
  p.W = c->W; p.adam = c->adam;
  p.X = c->X; p.y = c->y; p.perm = (const long long*)c->perm; p.num_train = c->num_train; p.perm_off = perm_off;
  p.t0 = t0; p.n_steps = n_steps;
  p.lr = c->lr; p.beta1 = c->beta1; p.beta2 = c->beta2; p.eps = c->eps; p.l2_reg = c->l2_reg; p.p_drop = c->p_drop;
  p.seed = c->seed;
  p.loss_out = loss_out; p.mask_in = mask_in;
  BNN_CUDA(cudaFuncSetAttribute(dropout_train_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  dropout_train_kernel<false><<<1, kDtThreads, plan.smem, (cudaStream_t)stream>>>(p);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t bnn_cdropout_train_supported(const BnnMlpDesc* d, int32_t batch) {
  if (!d) return 0;
  DtPlan plan;
  return dt_plan(d, batch, true, &plan) == 0;
}

extern "C" int64_t bnn_cdropout_train_u_per_step(const BnnMlpDesc* d, int32_t batch) {
  if (!d) { set_error("null desc"); return -1; }
  DtPlan plan;
  if (dt_plan(d, batch, true, &plan) < 0) return -1;
  return plan.p.mask_per_step;
}

extern "C" int32_t bnn_cdropout_train_run(const BnnCDropoutTrain* c, int64_t t0, int64_t perm_off, int32_t n_steps, int32_t epoch_end,
                                          float* stats_out, const float* u_in, void* stream) {
  if (!c || n_steps < 1 || t0 < 0 || perm_off < 0) BNN_FAIL("bnn_cdropout_train_run: bad arguments");
  if (!c->W || !c->adam || !c->p_logit || !c->p_adam || !c->noise_level || !c->X || !c->y || !c->perm || c->num_train < 1)
    BNN_FAIL("bnn_cdropout_train_run: W, adam, p_logit, p_adam, noise_level, X, y, perm and num_train are required");
  if (!(c->lr > 0.f)) BNN_FAIL("bnn_cdropout_train_run: lr must be positive");
  if (perm_off + (int64_t)(n_steps - 1) * c->batch >= c->num_train)
    BNN_FAIL("bnn_cdropout_train_run: %d steps of batch %d from offset %lld run past the epoch (%lld rows)", n_steps, c->batch,
             (long long)perm_off, (long long)c->num_train);
  DtPlan plan;
  if (dt_plan(&c->desc, c->batch, true, &plan)) return -1;
  DtParams& p = plan.p;
  p.W = c->W; p.adam = c->adam; p.plogit = c->p_logit; p.padam = c->p_adam; p.noise_level = c->noise_level;
  p.X = c->X; p.y = c->y; p.perm = (const long long*)c->perm; p.num_train = c->num_train; p.perm_off = perm_off;
  p.t0 = t0; p.n_steps = n_steps; p.epoch_end = epoch_end;
  p.lr = c->lr; p.beta1 = c->beta1; p.beta2 = c->beta2; p.eps = c->eps;
  p.lscale2 = c->lscale * c->lscale; p.dr = c->dr; p.min_noise = c->min_noise;
  p.seed = c->seed;
  p.stats_out = stats_out; p.mask_in = u_in;
  BNN_CUDA(cudaFuncSetAttribute(dropout_train_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  dropout_train_kernel<true><<<1, kDtThreads, plan.smem, (cudaStream_t)stream>>>(p);
  BNN_CUDA(cudaGetLastError());
  return 0;
}
