// Row f2 of SURVEY.md 8: the Bayes-by-Backprop minibatch training step of BNN_BBB.train (BNN_BBB.py:112-137) for networks that
// fit ONE SM -- the reference's own configuration (1-50-50-1, batch 32; cfg 2) is 2,852 packed parameters -- as a persistent
// single-CTA kernel that runs n_steps steps per launch:
//   minibatch gather (DataLoader batch, :116-120) -> L local-reparameterisation layers (GaussianLinear.forward, :29-37:
//   out = x mu^T + sqrt(x^2 s2^T) * eps, s2 = softplus(max(rho, -35)), eps in-kernel Philox) -> MSE (:104-106) -> backward
//   through the reparameterisation (what loss.backward() computes, :124) -> analytic KL(q || N(0, weight_std)) gradient
//   (:107-110, closed form, no autograd) -> Adam on (mu, rho) (torch.optim.Adam defaults, :114,125).
// The reference pays ~60 small kernels + two host synchronisations per step; here the variational parameters, their Adam
// moments, the activations and all gradients live in shared memory for the whole launch and a step is ~15 CTA barriers.
// Larger networks (9 packed vectors + activations beyond 227 KB of shared memory) are refused by bnn_bbb_train_supported and
// keep the plain-torch recipe in bnn_b200/BNN_BBB.py.
//
// Layout: mu / rho are the packed bias-last vectors of the forward path ([out_l][ld_l = roundup4(in_l + 1)] per layer, bias in
// column in_l, zero pad).  Activations are augmented rows (a | 1 | 0) so that the bias is one more column of every product.
#include <math.h>

#include "common.cuh"
#include "philox.cuh"

namespace bnn {

constexpr int kBtThreads = 1024;
constexpr uint32_t STREAM_BBB_LOCAL = 5;   // Philox stream of the local-reparameterisation noise: (seed, step) x unit index

struct BtLayer {
  int in, out, ld;
  long long off;      // float offset of the layer in the packed vectors
  int a_off;          // shared-memory float offset of this layer's input activations [B][ld]
  int s_off;          // ... of sqrt(var) [B][out] and, behind it, eps [B][out]
  long long e_off;    // offset of this layer's units in the per-step noise block ([B][out] each)
};

struct BtParams {
  int L, act, B, nout, D;
  BtLayer ly[BNN_MAX_LINEAR];
  long long stride;             // packed length (multiple of 4)
  long long eps_per_step;       // sum_l B * out_l
  float* mu; float* rho; float* adam;       // global: [stride], [stride], [4][stride] = m_mu, v_mu, m_rho, v_rho
  const float* X; const float* y; const long long* perm;
  long long num_train, perm_off, t0;
  int n_steps;
  float lr, beta1, beta2, eps, kl_factor, inv_w2;
  unsigned long long seed;
  float* mse_out; float* kl_out;            // [n_steps] or null: the step's mse / kl_factor-less KL, evaluated BEFORE the update
  const float* eps_in;                      // tests: injected noise [n_steps][eps_per_step] (layer-major, then [b][j]); null = Philox
  int sm_par, sm_act, sm_g0, sm_g1, sm_gv, sm_y, sm_red;   // shared-memory float offsets
  int max_w;                                // widest layer (outputs or inputs + 1)
};

__device__ __forceinline__ float bt_s2(float rho) {          // util.py:52-53: softplus(clamp(rho, min = -35)), torch threshold 20
  const float x = fmaxf(rho, -35.0f);
  return x > 20.0f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float bt_ds2(float rho) {         // d s2 / d rho (clamp passes no gradient below -35)
  if (rho < -35.0f) return 0.0f;
  if (rho > 20.0f) return 1.0f;
  return 1.0f / (1.0f + expf(-rho));
}
__device__ __forceinline__ float bt_act_bwd(float a, int act) {   // derivative through the activation's OUTPUT
  switch (act) {
    case BNN_ACT_RELU: return a > 0.0f ? 1.0f : 0.0f;
    case BNN_ACT_TANH: return 1.0f - a * a;
    case BNN_ACT_SIGMOID: return a * (1.0f - a);
    default: return 1.0f;
  }
}

__global__ void __launch_bounds__(kBtThreads, 1) bbb_train_kernel(const __grid_constant__ BtParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const long long P = p.stride;
  float* s_mu = sm + p.sm_par;
  float* s_rho = s_mu + P;
  float* s_s2 = s_rho + P;
  float* s_mm = s_s2 + P;      // Adam: m_mu, v_mu, m_rho, v_rho
  float* s_vm = s_mm + P;
  float* s_mr = s_vm + P;
  float* s_vr = s_mr + P;
  float* s_gmu = s_vr + P;     // data-term gradients of the step
  float* s_gs2 = s_gmu + P;
  float* s_act = sm + p.sm_act;
  float* s_g0 = sm + p.sm_g0;  // dL/d(out) of the current layer [B][max_w]
  float* s_g1 = sm + p.sm_g1;  // dL/d(input activations) -> next g0
  float* s_gv = sm + p.sm_gv;  // dL/d(var) of the current layer [B][max_w]
  float* s_y = sm + p.sm_y;    // targets [B][nout]
  double* s_red = reinterpret_cast<double*>(sm + p.sm_red);   // [8 warps][2]: mse / kl partial sums

  for (long long i = tid; i < P; i += kBtThreads) {
    const float r = p.rho[i];
    s_mu[i] = p.mu[i]; s_rho[i] = r; s_s2[i] = bt_s2(r);
    s_mm[i] = p.adam[i]; s_vm[i] = p.adam[P + i]; s_mr[i] = p.adam[2 * P + i]; s_vr[i] = p.adam[3 * P + i];
  }
  __syncthreads();
  double b1t = pow((double)p.beta1, (double)p.t0), b2t = pow((double)p.beta2, (double)p.t0);   // beta^t so far (every thread: cheap, once)
  const int L = p.L, B = p.B;

#pragma unroll 1
  for (int it = 0; it < p.n_steps; ++it) {
    const long long boff = p.perm_off + (long long)it * B;
    const long long left = p.num_train - boff;
    const int rows = left < B ? (int)left : B;                     // the last batch of an epoch may be short
    // ---- gather: augmented rows (x | 1 | 0) of the batch, and the targets
    {
      const BtLayer& Y0 = p.ly[0];
      float* a0 = s_act + Y0.a_off;
      for (int i = tid; i < B * Y0.ld; i += kBtThreads) {
        const int b = i / Y0.ld, k = i - b * Y0.ld;
        float v = 0.0f;
        if (b < rows) {
          if (k < p.D) v = __ldg(p.X + __ldg(p.perm + boff + b) * p.D + k);
          else if (k == p.D) v = 1.0f;
        }
        a0[i] = v;
      }
      for (int i = tid; i < B * p.nout; i += kBtThreads) {
        const int b = i / p.nout;
        s_y[i] = b < rows ? __ldg(p.y + __ldg(p.perm + boff + b) * p.nout + (i - b * p.nout)) : 0.0f;
      }
      // the step's local-reparameterisation noise for ALL layers, four normals per Philox block over the step's flat noise index
      // (layer-major, then [b][j]) -- drawn here, under the latency of the gather's global loads, instead of one block (and one
      // Box-Muller pair) per unit inside the forward
      if (!p.eps_in) {
        const long long nq = (p.eps_per_step + 3) >> 2;
        for (long long q = tid; q < nq; q += kBtThreads) {
          float z[4];
          philox_normal4(p.seed, (uint32_t)(p.t0 + it), STREAM_BBB_LOCAL, (unsigned long long)q, z);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const long long u = q * 4 + c;
            if (u < p.eps_per_step) {
              int l = 0;
              while (l + 1 < L && u >= p.ly[l + 1].e_off) ++l;
              s_act[p.ly[l].s_off + B * p.ly[l].out + (int)(u - p.ly[l].e_off)] = z[c];
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- forward: local reparameterisation, layer by layer (thread = one output unit of one row)
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const BtLayer& Y = p.ly[l];
      const float* a = s_act + Y.a_off;
      float* sd = s_act + Y.s_off;
      float* ep = sd + B * Y.out;
      float* nxt = l + 1 < L ? s_act + p.ly[l + 1].a_off : s_g1;   // the last layer's output goes to s_g1 as "pred"
      const int ldn = l + 1 < L ? p.ly[l + 1].ld : p.nout;
      for (int i = tid; i < B * Y.out; i += kBtThreads) {
        const int b = i / Y.out, j = i - b * Y.out;
        const float* ar = a + b * Y.ld;
        const float* mr = s_mu + Y.off + (long long)j * Y.ld;
        const float* sr = s_s2 + Y.off + (long long)j * Y.ld;
        // rows are 16-byte aligned and zero-padded to ld = roundup4(in + 1): float4 along k (conflict-free: the row pitch of
        // consecutive output units is an odd number of 16-byte groups for the widths that matter)
        float mean = 0.0f, var = 0.0f;
        for (int k = 0; k < Y.ld; k += 4) {
          const float4 x = *reinterpret_cast<const float4*>(ar + k);
          const float4 m = *reinterpret_cast<const float4*>(mr + k);
          const float4 v = *reinterpret_cast<const float4*>(sr + k);
          mean = fmaf(x.x, m.x, mean); mean = fmaf(x.y, m.y, mean); mean = fmaf(x.z, m.z, mean); mean = fmaf(x.w, m.w, mean);
          var = fmaf(x.x * x.x, v.x, var); var = fmaf(x.y * x.y, v.y, var); var = fmaf(x.z * x.z, v.z, var); var = fmaf(x.w * x.w, v.w, var);
        }
        const float e = p.eps_in ? p.eps_in[(long long)it * p.eps_per_step + Y.e_off + i] : ep[i];
        const float s = sqrtf(var);
        sd[i] = s; ep[i] = e;
        const float o = b < rows ? mean + s * e : 0.0f;
        nxt[b * ldn + j] = l + 1 < L ? apply_act(o, p.act) : o;
      }
      if (l + 1 < L) {      // ones column and zero pad of the next layer's augmented input
        const BtLayer& N = p.ly[l + 1];
        float* an = s_act + N.a_off;
        for (int i = tid; i < B * (N.ld - N.in); i += kBtThreads) {
          const int b = i / (N.ld - N.in), k = N.in + (i - b * (N.ld - N.in));
          an[b * N.ld + k] = (k == N.in && b < rows) ? 1.0f : 0.0f;
        }
      }
      __syncthreads();
    }
    // ---- loss: mse = mean((pred - y)^2) over rows x nout (MSELoss(reduction='mean')); dL/d(pred) -> s_g0
    {
      const float inv_n = 1.0f / (float)(rows * p.nout);
      double part = 0.0;
      for (int i = tid; i < B * p.nout; i += kBtThreads) {
        const int b = i / p.nout;
        const float r = b < rows ? s_g1[i] - s_y[i] : 0.0f;
        part += (double)r * (double)r;
        s_g0[b * p.max_w + (i - b * p.nout)] = 2.0f * r * inv_n;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if ((tid & 31) == 0) s_red[(tid >> 5) * 2] = part;
    }
    __syncthreads();
    // ---- backward through the layers: data-term gradients of (mu, s2) and of the layer input
#pragma unroll 1
    for (int l = L - 1; l >= 0; --l) {
      const BtLayer& Y = p.ly[l];
      const float* a = s_act + Y.a_off;
      const float* sd = s_act + Y.s_off;
      const float* ep = sd + B * Y.out;
      // dL/d(var) = dL/d(out) * eps / (2 sqrt(var))
      for (int i = tid; i < B * Y.out; i += kBtThreads) {
        const int b = i / Y.out, j = i - b * Y.out;
        const float s = sd[i];                                       // > 0 on live rows (the bias term alone is s2 > 0); 0 on dead rows
        s_gv[b * p.max_w + j] = s > 0.0f ? s_g0[b * p.max_w + j] * ep[i] / (2.0f * s) : 0.0f;
      }
      __syncthreads();
      // parameters: thread = one (j, k) entry, sum over the batch
      // parameters: thread = four consecutive (j, k..k+3) entries, sum over the batch (pad columns come out as zero)
      {
        const int q4 = Y.ld >> 2;
        for (int i = tid; i < Y.out * q4; i += kBtThreads) {
          const int j = i / q4, k = (i - j * q4) * 4;
          float4 gm = make_float4(0.f, 0.f, 0.f, 0.f), gs = gm;
          for (int b = 0; b < B; ++b) {
            const float4 x = *reinterpret_cast<const float4*>(a + b * Y.ld + k);
            const float g0 = s_g0[b * p.max_w + j], gv = s_gv[b * p.max_w + j];
            gm.x = fmaf(g0, x.x, gm.x); gm.y = fmaf(g0, x.y, gm.y); gm.z = fmaf(g0, x.z, gm.z); gm.w = fmaf(g0, x.w, gm.w);
            gs.x = fmaf(gv, x.x * x.x, gs.x); gs.y = fmaf(gv, x.y * x.y, gs.y); gs.z = fmaf(gv, x.z * x.z, gs.z); gs.w = fmaf(gv, x.w * x.w, gs.w);
          }
          *reinterpret_cast<float4*>(s_gmu + Y.off + (long long)j * Y.ld + k) = gm;
          *reinterpret_cast<float4*>(s_gs2 + Y.off + (long long)j * Y.ld + k) = gs;
        }
      }
      // layer input: thread = one (b, k) entry, sum over the outputs; then through the activation
      if (l > 0) {
        const int q4 = (Y.in + 3) >> 2;
        for (int i = tid; i < B * q4; i += kBtThreads) {
          const int b = i / q4, k = (i - b * q4) * 4;
          const float4 x = *reinterpret_cast<const float4*>(a + b * Y.ld + k);
          float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int j = 0; j < Y.out; ++j) {
            const long long w = Y.off + (long long)j * Y.ld + k;
            const float4 m = *reinterpret_cast<const float4*>(s_mu + w);
            const float4 v = *reinterpret_cast<const float4*>(s_s2 + w);
            const float g0 = s_g0[b * p.max_w + j], gv2 = 2.0f * s_gv[b * p.max_w + j];
            g.x = fmaf(g0, m.x, g.x); g.y = fmaf(g0, m.y, g.y); g.z = fmaf(g0, m.z, g.z); g.w = fmaf(g0, m.w, g.w);
            g.x = fmaf(gv2 * x.x, v.x, g.x); g.y = fmaf(gv2 * x.y, v.y, g.y); g.z = fmaf(gv2 * x.z, v.z, g.z); g.w = fmaf(gv2 * x.w, v.w, g.w);
          }
          const float xs[4] = {x.x, x.y, x.z, x.w}, gs4[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (k + c < Y.in) s_g1[b * p.max_w + k + c] = gs4[c] * bt_act_bwd(xs[c], p.act);
        }
      }
      __syncthreads();
      if (l > 0) {
        for (int i = tid; i < B * Y.in; i += kBtThreads) {
          const int b = i / Y.in, k = i - b * Y.in;
          s_g0[b * p.max_w + k] = s_g1[b * p.max_w + k];
        }
        __syncthreads();
      }
    }
    // ---- KL gradient (closed form) + Adam on (mu, rho); KL value of the step's input parameters
    {
      b1t *= (double)p.beta1; b2t *= (double)p.beta2;
      const float step_size = (float)((double)p.lr / (1.0 - b1t));
      const float bc2_sqrt = (float)sqrt(1.0 - b2t);
      const float om1 = 1.0f - p.beta1, om2 = 1.0f - p.beta2;
      double klp = 0.0;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const BtLayer& Y = p.ly[l];
        for (int i = tid; i < Y.out * (Y.in + 1); i += kBtThreads) {
          const int j = i / (Y.in + 1), k = i - j * (Y.in + 1);
          const long long w = Y.off + (long long)j * Y.ld + k;
          const float mu = s_mu[w], rho = s_rho[w], s2 = s_s2[w];
          // KL(N(mu, s) || N(0, w0)) = 0.5 (s2 / w0^2 + mu^2 / w0^2 - 1 - log(s2 / w0^2))          (BNN_BBB.py:107-110)
          if (p.kl_out) {
            const float vr = s2 * p.inv_w2;
            klp += 0.5 * ((double)vr + (double)(mu * mu * p.inv_w2) - 1.0 - (double)logf(vr));
          }
          const float g_mu = s_gmu[w] + p.kl_factor * mu * p.inv_w2;
          const float g_rho = (s_gs2[w] + p.kl_factor * 0.5f * (p.inv_w2 - 1.0f / s2)) * bt_ds2(rho);
          float m = s_mm[w], v = s_vm[w];
          m = p.beta1 * m + om1 * g_mu; v = p.beta2 * v + om2 * g_mu * g_mu;
          s_mm[w] = m; s_vm[w] = v;
          s_mu[w] = mu - step_size * (m / (sqrtf(v) / bc2_sqrt + p.eps));
          m = s_mr[w]; v = s_vr[w];
          m = p.beta1 * m + om1 * g_rho; v = p.beta2 * v + om2 * g_rho * g_rho;
          s_mr[w] = m; s_vr[w] = v;
          const float rn = rho - step_size * (m / (sqrtf(v) / bc2_sqrt + p.eps));
          s_rho[w] = rn; s_s2[w] = bt_s2(rn);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) klp += __shfl_xor_sync(0xffffffffu, klp, o);
      if ((tid & 31) == 0) s_red[(tid >> 5) * 2 + 1] = klp;
    }
    __syncthreads();
    if (tid == 0) {
      double mse = 0.0, kl = 0.0;
      for (int w = 0; w < kBtThreads / 32; ++w) { mse += s_red[w * 2]; kl += s_red[w * 2 + 1]; }
      if (p.mse_out) p.mse_out[it] = (float)(mse / (double)(rows * p.nout));
      if (p.kl_out) p.kl_out[it] = (float)kl;
    }
    __syncthreads();
  }
  for (long long i = tid; i < P; i += kBtThreads) {
    p.mu[i] = s_mu[i]; p.rho[i] = s_rho[i];
    p.adam[i] = s_mm[i]; p.adam[P + i] = s_vm[i]; p.adam[2 * P + i] = s_mr[i]; p.adam[3 * P + i] = s_vr[i];
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
struct BtPlan {
  BtParams p;
  size_t smem;
};

// 0 = ok, 1 = does not fit one SM (message set), < 0 = bad arguments
static int bt_plan(const BnnMlpDesc* d, int batch, BtPlan* plan) {
  Layout Lo;
  if (make_layout(d, &Lo)) return -1;
  if (batch < 1 || batch > 1024) { set_error("bnn_bbb_train: batch must be in [1, 1024] (got %d)", batch); return -1; }
  BtParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.L = Lo.n_lin; p.act = Lo.act; p.B = batch; p.nout = Lo.out[Lo.n_lin - 1]; p.D = Lo.in[0];
  p.stride = (Lo.off[Lo.n_lin] + 3) & ~3LL;
  long long f = 0, e = 0;
  int max_w = p.nout;
  long long act = 0;
  for (int l = 0; l < p.L; ++l) {
    BtLayer& Y = p.ly[l];
    Y.in = Lo.in[l]; Y.out = Lo.out[l]; Y.ld = Lo.ld[l]; Y.off = Lo.off[l];
    Y.a_off = (int)act; act += (long long)batch * Y.ld;            // multiple of 4: float4 rows
    Y.s_off = (int)act; act += (2LL * batch * Y.out + 3) & ~3LL;
    Y.e_off = e; e += (long long)batch * Y.out;
    if (Y.out > max_w) max_w = Y.out;
    if (Y.in + 1 > max_w) max_w = Y.in + 1;
  }
  p.eps_per_step = e;
  p.max_w = max_w;
  p.sm_par = 0; f = 9 * p.stride;
  p.sm_act = (int)f; f += (act + 3) & ~3LL;
  p.sm_g0 = (int)f; f += ((long long)batch * max_w + 3) & ~3LL;
  p.sm_g1 = (int)f; f += ((long long)batch * max_w + 3) & ~3LL;
  p.sm_gv = (int)f; f += ((long long)batch * max_w + 3) & ~3LL;
  p.sm_y = (int)f; f += ((long long)batch * p.nout + 3) & ~3LL;
  p.sm_red = (int)f; f += 2 * 2 * (kBtThreads / 32);
  plan->smem = (size_t)f * 4;
  if (plan->smem > (size_t)227 * 1024) {
    set_error("bnn_bbb_train: %lld packed parameters at batch %d need %zu bytes of shared memory (one SM has 227 KB)",
              (long long)p.stride, batch, plan->smem);
    return 1;
  }
  return 0;
}

}  // namespace bnn

using namespace bnn;

extern "C" int32_t bnn_bbb_train_supported(const BnnMlpDesc* d, int32_t batch) {
  if (!d) return 0;
  BtPlan plan;
  return bt_plan(d, batch, &plan) == 0;
}

extern "C" int64_t bnn_bbb_train_eps_per_step(const BnnMlpDesc* d, int32_t batch) {
  if (!d) { set_error("null desc"); return -1; }
  BtPlan plan;
  if (bt_plan(d, batch, &plan) < 0) return -1;
  return plan.p.eps_per_step;
}

extern "C" int32_t bnn_bbb_train_run(const BnnBbbTrain* c, int64_t t0, int64_t perm_off, int32_t n_steps, float* mse_out,
                                     float* kl_out, const float* eps_in, void* stream) {
  if (!c || n_steps < 1 || t0 < 0 || perm_off < 0) BNN_FAIL("bnn_bbb_train_run: bad arguments");
  if (!c->mu || !c->rho || !c->adam || !c->X || !c->y || !c->perm || c->num_train < 1)
    BNN_FAIL("bnn_bbb_train_run: mu, rho, adam, X, y, perm and num_train are required");
  if (!(c->weight_std > 0.f) || !(c->lr > 0.f)) BNN_FAIL("bnn_bbb_train_run: weight_std and lr must be positive");
  if (perm_off + (int64_t)(n_steps - 1) * c->batch >= c->num_train)
    BNN_FAIL("bnn_bbb_train_run: %d steps of batch %d from offset %lld run past the epoch (%lld rows)", n_steps, c->batch,
             (long long)perm_off, (long long)c->num_train);
  BtPlan plan;
  const int rc = bt_plan(&c->desc, c->batch, &plan);
  if (rc) return -1;
  BtParams& p = plan.p;
  p.mu = c->mu; p.rho = c->rho; p.adam = c->adam;
  p.X = c->X; p.y = c->y; p.perm = (const long long*)c->perm; p.num_train = c->num_train; p.perm_off = perm_off;
  p.t0 = t0; p.n_steps = n_steps;
  p.lr = c->lr; p.beta1 = c->beta1; p.beta2 = c->beta2; p.eps = c->eps; p.kl_factor = c->kl_factor;
  p.inv_w2 = 1.0f / (c->weight_std * c->weight_std);
  p.seed = c->seed;
  p.mse_out = mse_out; p.kl_out = kl_out; p.eps_in = eps_in;
  BNN_CUDA(cudaFuncSetAttribute(bbb_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  bbb_train_kernel<<<1, kBtThreads, plan.smem, (cudaStream_t)stream>>>(p);
  BNN_CUDA(cudaGetLastError());
  return 0;
}
