// Row f2 of SURVEY.md 8, second half: the MC-dropout minibatch training step of BNN_Dropout.train (BNN_Dropout.py:39-65) for
// networks that fit ONE SM, as a persistent single-CTA kernel that runs n_steps steps per launch:
//   minibatch gather (:53-56) -> NN_Dropout.forward (:15-21: every Linear whose input is wider than one unit sees its input
//   multiplied by a fresh Bernoulli(1 - p) keep mask; F.dropout(x, p) * (1 - p)) -> MSELoss(reduction='sum') (:55,59) ->
//   backward -> Adam with weight decay l2_reg on the weights, none on the biases (:43-52,61).
// Same structure as csrc/bbb_train.cu: the packed network (bias-last rows, the layout of the forward path), its two Adam moment
// vectors, the masked activations and two gradient panels live in shared memory for the whole launch.  The keep masks are never
// stored: forward and backward evaluate the same counter-based draw (Philox stream 6, subsequence = step) or read the injected
// mask (tests: the reference's own F.dropout draws).  The parameter gradient is consumed by Adam in the thread that forms it.
// The reference's cfg 3 network (9-100-100-1, batch 32: 11,704 packed parameters) needs 197 KB.
#include <math.h>

#include "common.cuh"
#include "philox.cuh"

namespace bnn {

constexpr int kDtThreads = 1024;
constexpr uint32_t STREAM_DROPOUT_TRAIN = 6;

struct DtLayer {
  int in, out, ld, masked;
  long long off;      // float offset of the layer in the packed vector
  int a_off;          // shared-memory float offset of this layer's (masked, augmented) input [B][ld]
  long long m_off;    // offset of this layer's keep mask in the per-step mask block ([B][in]); -1 = not masked
};

struct DtParams {
  int L, act, B, nout, D;
  DtLayer ly[BNN_MAX_LINEAR];
  long long stride, mask_per_step;
  float* W; float* adam;                    // global: [stride], [2][stride] = m, v
  const float* X; const float* y; const long long* perm;
  long long num_train, perm_off, t0;
  int n_steps;
  float lr, beta1, beta2, eps, l2_reg, p_drop;
  unsigned long long seed;
  float* loss_out;                          // [n_steps] or null: sum of squared errors of the step's batch, before its update
  const float* mask_in;                     // tests: injected keep factors [n_steps][mask_per_step]; null = Philox
  int sm_act, sm_g0, sm_g1, sm_y, sm_red, max_w;
};

// keep factor of input unit k of row b of layer l at launch step `it`
__device__ __forceinline__ float dt_keep(const DtParams& p, const DtLayer& Y, int it, int b, int k) {
  const long long u = Y.m_off + (long long)b * Y.in + k;
  if (p.mask_in) return p.mask_in[(long long)it * p.mask_per_step + u];
  const u32x4 w = philox_block(p.seed, (uint32_t)(p.t0 + it), STREAM_DROPOUT_TRAIN, (uint64_t)(u >> 2));
  return bernoulli_keep(u24(pick_word(w, (int)(u & 3))), p.p_drop);
}
__device__ __forceinline__ float dt_act_bwd(float a, int act) {   // derivative through the activation's OUTPUT
  switch (act) {
    case BNN_ACT_RELU: return a > 0.0f ? 1.0f : 0.0f;
    case BNN_ACT_TANH: return 1.0f - a * a;
    case BNN_ACT_SIGMOID: return a * (1.0f - a);
    default: return 1.0f;
  }
}

__global__ void __launch_bounds__(kDtThreads, 1) dropout_train_kernel(const __grid_constant__ DtParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const long long P = p.stride;
  float* s_w = sm;
  float* s_m = s_w + P;
  float* s_v = s_m + P;
  float* s_act = sm + p.sm_act;
  float* s_ga = sm + p.sm_g0;   // dL/d(out) of the current layer [B][max_w] ...
  float* s_gb = sm + p.sm_g1;   // ... and of the layer below (ping-pong)
  float* s_y = sm + p.sm_y;
  double* s_red = reinterpret_cast<double*>(sm + p.sm_red);   // [warps]: loss partial sums

  for (long long i = tid; i < P; i += kDtThreads) { s_w[i] = p.W[i]; s_m[i] = p.adam[i]; s_v[i] = p.adam[P + i]; }
  __syncthreads();
  double b1t = pow((double)p.beta1, (double)p.t0), b2t = pow((double)p.beta2, (double)p.t0);
  const int L = p.L, B = p.B;

#pragma unroll 1
  for (int it = 0; it < p.n_steps; ++it) {
    const long long boff = p.perm_off + (long long)it * B;
    const long long left = p.num_train - boff;
    const int rows = left < B ? (int)left : B;                     // the last batch of an epoch may be short
    // ---- gather: masked, augmented rows (x * keep | 1 | 0) of the batch, and the targets
    {
      const DtLayer& Y0 = p.ly[0];
      float* a0 = s_act + Y0.a_off;
      for (int i = tid; i < B * Y0.ld; i += kDtThreads) {
        const int b = i / Y0.ld, k = i - b * Y0.ld;
        float v = 0.0f;
        if (b < rows) {
          if (k < p.D) {
            v = __ldg(p.X + __ldg(p.perm + boff + b) * p.D + k);
            if (Y0.masked) v *= dt_keep(p, Y0, it, b, k);
          } else if (k == p.D) v = 1.0f;
        }
        a0[i] = v;
      }
      for (int i = tid; i < B * p.nout; i += kDtThreads) {
        const int b = i / p.nout;
        s_y[i] = b < rows ? __ldg(p.y + __ldg(p.perm + boff + b) * p.nout + (i - b * p.nout)) : 0.0f;
      }
    }
    __syncthreads();
    // ---- forward (thread = one output unit of one row; float4 along k over 16-byte aligned, zero-padded rows)
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const DtLayer& Y = p.ly[l];
      const float* a = s_act + Y.a_off;
      const bool last = l + 1 == L;
      float* nxt = last ? s_gb : s_act + p.ly[l + 1].a_off;        // the last layer's output goes to s_gb as "pred"
      const int ldn = last ? p.nout : p.ly[l + 1].ld;
      for (int i = tid; i < B * Y.out; i += kDtThreads) {
        const int b = i / Y.out, j = i - b * Y.out;
        const float* ar = a + b * Y.ld;
        const float* wr = s_w + Y.off + (long long)j * Y.ld;
        float o = 0.0f;
        for (int k = 0; k < Y.ld; k += 4) {
          const float4 x = *reinterpret_cast<const float4*>(ar + k);
          const float4 w = *reinterpret_cast<const float4*>(wr + k);
          o = fmaf(x.x, w.x, o); o = fmaf(x.y, w.y, o); o = fmaf(x.z, w.z, o); o = fmaf(x.w, w.w, o);
        }
        if (b >= rows) o = 0.0f;
        if (!last) {
          o = apply_act(o, p.act);
          const DtLayer& N = p.ly[l + 1];
          if (N.masked && b < rows) o *= dt_keep(p, N, it, b, j);
        }
        nxt[b * ldn + j] = o;
      }
      if (!last) {          // ones column and zero pad of the next layer's augmented input
        const DtLayer& N = p.ly[l + 1];
        float* an = s_act + N.a_off;
        for (int i = tid; i < B * (N.ld - N.in); i += kDtThreads) {
          const int b = i / (N.ld - N.in), k = N.in + (i - b * (N.ld - N.in));
          an[b * N.ld + k] = (k == N.in && b < rows) ? 1.0f : 0.0f;
        }
      }
      __syncthreads();
    }
    // ---- loss = sum (pred - y)^2 (MSELoss(reduction='sum'), BNN_Dropout.py:55); dL/d(pred) -> s_ga
    {
      double part = 0.0;
      for (int i = tid; i < B * p.nout; i += kDtThreads) {
        const int b = i / p.nout;
        const float r = b < rows ? s_gb[i] - s_y[i] : 0.0f;
        part += (double)r * (double)r;
        s_ga[b * p.max_w + (i - b * p.nout)] = 2.0f * r;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if ((tid & 31) == 0) s_red[tid >> 5] = part;
    }
    __syncthreads();
    if (tid == 0 && p.loss_out) {
      double s = 0.0;
      for (int w = 0; w < kDtThreads / 32; ++w) s += s_red[w];
      p.loss_out[it] = (float)s;
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
      // weights BEFORE their update
      if (l > 0) {
        const int q4 = (Y.in + 3) >> 2;
        for (int i = tid; i < B * q4; i += kDtThreads) {
          const int b = i / q4, k = (i - b * q4) * 4;
          float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int j = 0; j < Y.out; ++j) {
            const float4 w = *reinterpret_cast<const float4*>(s_w + Y.off + (long long)j * Y.ld + k);
            const float gj = g0[b * p.max_w + j];
            g.x = fmaf(gj, w.x, g.x); g.y = fmaf(gj, w.y, g.y); g.z = fmaf(gj, w.z, g.z); g.w = fmaf(gj, w.w, g.w);
          }
          const float4 x = *reinterpret_cast<const float4*>(a + b * Y.ld + k);   // act(z) * keep
          const float xs[4] = {x.x, x.y, x.z, x.w}, gs[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (k + c < Y.in) {
              float keep = 1.0f;
              if (Y.masked) keep = b < rows ? dt_keep(p, Y, it, b, k + c) : 0.0f;
              // keep is 0 or (within an ulp of) 1: where it is not 0, the stored value is the activation's output itself
              g1[b * p.max_w + k + c] = keep != 0.0f ? gs[c] * keep * dt_act_bwd(xs[c] / keep, p.act) : 0.0f;
            }
          }
        }
      }
      __syncthreads();
      // parameters: thread = four consecutive (j, k..k+3) entries: gradient over the batch, weight decay, Adam, in place
      {
        const int q4 = Y.ld >> 2;
        for (int i = tid; i < Y.out * q4; i += kDtThreads) {
          const int j = i / q4, k = (i - j * q4) * 4;
          float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int b = 0; b < B; ++b) {
            const float4 x = *reinterpret_cast<const float4*>(a + b * Y.ld + k);
            const float gj = g0[b * p.max_w + j];
            g.x = fmaf(gj, x.x, g.x); g.y = fmaf(gj, x.y, g.y); g.z = fmaf(gj, x.z, g.z); g.w = fmaf(gj, x.w, g.w);
          }
          const float gs[4] = {g.x, g.y, g.z, g.w};
          const long long w0 = Y.off + (long long)j * Y.ld + k;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (k + c <= Y.in) {
              const long long w = w0 + c;
              const float wv = s_w[w];
              const float gr = k + c < Y.in ? fmaf(p.l2_reg, wv, gs[c]) : gs[c];   // Adam's weight_decay: weights only (:43-51)
              const float m = p.beta1 * s_m[w] + om1 * gr, v = p.beta2 * s_v[w] + om2 * gr * gr;
              s_m[w] = m; s_v[w] = v;
              s_w[w] = wv - step_size * (m / (sqrtf(v) / bc2_sqrt + p.eps));
            }
          }
        }
      }
      __syncthreads();
      float* t = g0; g0 = g1; g1 = t;
    }
  }
  for (long long i = tid; i < P; i += kDtThreads) { p.W[i] = s_w[i]; p.adam[i] = s_m[i]; p.adam[P + i] = s_v[i]; }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
struct DtPlan {
  DtParams p;
  size_t smem;
};

// 0 = ok, 1 = does not fit one SM (message set), < 0 = bad arguments
static int dt_plan(const BnnMlpDesc* d, int batch, DtPlan* plan) {
  Layout Lo;
  if (make_layout(d, &Lo)) return -1;
  if (batch < 1 || batch > 1024) { set_error("bnn_dropout_train: batch must be in [1, 1024] (got %d)", batch); return -1; }
  DtParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.L = Lo.n_lin; p.act = Lo.act; p.B = batch; p.nout = Lo.out[Lo.n_lin - 1]; p.D = Lo.in[0];
  p.stride = (Lo.off[Lo.n_lin] + 3) & ~3LL;
  long long act = 0, m = 0;
  int max_w = p.nout;
  for (int l = 0; l < p.L; ++l) {
    DtLayer& Y = p.ly[l];
    Y.in = Lo.in[l]; Y.out = Lo.out[l]; Y.ld = Lo.ld[l]; Y.off = Lo.off[l];
    Y.masked = Y.in > 1 ? 1 : 0;                                   // BNN_Dropout.py:18: "x.shape[1] > 1"
    Y.a_off = (int)act; act += (long long)batch * Y.ld;
    Y.m_off = Y.masked ? m : -1;
    if (Y.masked) m += ((long long)batch * Y.in + 3) & ~3LL;
    if (Y.out > max_w) max_w = Y.out;
    if (Y.in + 1 > max_w) max_w = Y.in + 1;
  }
  p.mask_per_step = m;
  p.max_w = max_w;
  long long f = 3 * p.stride;
  p.sm_act = (int)f; f += (act + 3) & ~3LL;
  p.sm_g0 = (int)f; f += ((long long)batch * max_w + 3) & ~3LL;
  p.sm_g1 = (int)f; f += ((long long)batch * max_w + 3) & ~3LL;
  p.sm_y = (int)f; f += ((long long)batch * p.nout + 3) & ~3LL;
  p.sm_red = (int)f; f += 2 * (kDtThreads / 32);
  plan->smem = (size_t)f * 4;
  if (plan->smem > (size_t)227 * 1024) {
    set_error("bnn_dropout_train: %lld packed parameters at batch %d need %zu bytes of shared memory (one SM has 227 KB)",
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
  return dt_plan(d, batch, &plan) == 0;
}

extern "C" int64_t bnn_dropout_train_mask_per_step(const BnnMlpDesc* d, int32_t batch) {
  if (!d) { set_error("null desc"); return -1; }
  DtPlan plan;
  if (dt_plan(d, batch, &plan) < 0) return -1;
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
  if (dt_plan(&c->desc, c->batch, &plan)) return -1;
  DtParams& p = plan.p;
  p.W = c->W; p.adam = c->adam;
  p.X = c->X; p.y = c->y; p.perm = (const long long*)c->perm; p.num_train = c->num_train; p.perm_off = perm_off;
  p.t0 = t0; p.n_steps = n_steps;
  p.lr = c->lr; p.beta1 = c->beta1; p.beta2 = c->beta2; p.eps = c->eps; p.l2_reg = c->l2_reg; p.p_drop = c->p_drop;
  p.seed = c->seed;
  p.loss_out = loss_out; p.mask_in = mask_in;
  BNN_CUDA(cudaFuncSetAttribute(dropout_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  dropout_train_kernel<<<1, kDtThreads, plan.smem, (cudaStream_t)stream>>>(p);
  BNN_CUDA(cudaGetLastError());
  return 0;
}
