// Kernel (e): predictive moments and metrics.
//   * finalize / merge: Chan-merge of per-group (or per-GPU) (count, mean, M2) in FP64
//     -> mean, unbiased var (BNN.py:37 `preds.mean(0), preds.var(0)`; torch's CPU var
//     is a double-accumulated Welford, so FP64 state is what matches it);
//   * standalone moments over a materialised pred[S, M] (HBM-bound: 4*S*M bytes read);
//   * rmse / nll of BNN.validate (BNN.py:30-31) and of experiments/SGDMC.py:83-93.
#include "common.cuh"

namespace bnn {

// ---- finalize: part[G][2][M] -> mean/var (+ FP64 shard moments) --------------
__global__ void finalize_kernel(const double* __restrict__ part, int G, int Sg, long long S, long long M,
                                const long long* __restrict__ counts, float* __restrict__ mean, float* __restrict__ var,
                                double* __restrict__ out_mean, double* __restrict__ out_m2) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += stride) {
    double n = 0.0, mu = 0.0, m2 = 0.0;
    for (int g = 0; g < G; ++g) {
      double nb;
      if (counts) nb = (double)counts[g];
      else {
        const long long rem = S - (long long)g * Sg;
        nb = (double)(rem < Sg ? rem : Sg);
      }
      if (nb <= 0.0) continue;
      const double mb = part[(long long)g * 2 * M + m];
      const double m2b = part[(long long)g * 2 * M + M + m];
      const double d = mb - mu;
      const double nn = n + nb;
      mu += d * (nb / nn);
      m2 += m2b + d * d * (n * nb / nn);
      n = nn;
    }
    if (mean) mean[m] = (float)mu;
    if (var) var[m] = (float)(m2 / (n - 1.0));  // n == 1 -> 0/0 = NaN, as torch.var
    if (out_mean) out_mean[m] = mu;
    if (out_m2) out_m2[m] = m2;
  }
}

static int grid_for(long long work, int block) {
  long long b = (work + block - 1) / block;
  const long long cap = 8LL * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int finalize_moments(const double* part, const int* /*unused*/, int G, int Sg, long long S, long long M, float* mean, float* var,
                     double* part_mean, double* part_m2, cudaStream_t st) {
  finalize_kernel<<<grid_for(M, 256), 256, 0, st>>>(part, G, Sg, S, M, nullptr, mean, var, part_mean, part_m2);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- standalone moments over pred[S, M] ---------------------------------------
// thread = 4 consecutive columns x one group of rows; FP64 Welford per column.
constexpr int kMomBlock = 256;

__global__ void __launch_bounds__(kMomBlock) moments_kernel(const float* __restrict__ pred, long long S, long long M, int Sg,
                                                            double* __restrict__ part) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // column quad
  const long long m0 = q * 4;
  if (m0 >= M) return;
  const int g = blockIdx.y;
  const long long s0 = (long long)g * Sg;
  const long long s1 = s0 + Sg < S ? s0 + Sg : S;
  const bool vec = ((M & 3) == 0);  // rows stay 16B aligned
  const int nc = (M - m0) < 4 ? (int)(M - m0) : 4;
  double mu[4] = {0, 0, 0, 0}, m2[4] = {0, 0, 0, 0};
  long long s = s0;
  auto upd = [&](const float (&v)[4], double inv) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double d = (double)v[c] - mu[c];
      mu[c] += d * inv;
      m2[c] += d * ((double)v[c] - mu[c]);
    }
  };
  if (vec) {
    // 4 independent 16-byte loads in flight per thread
    for (; s + 4 <= s1; s += 4) {
      float4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = __ldcs(reinterpret_cast<const float4*>(pred + (s + u) * M + m0));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float v[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
        upd(v, 1.0 / (double)(s + u - s0 + 1));
      }
    }
  }
  for (; s < s1; ++s) {
    float v[4] = {0, 0, 0, 0};
    for (int c = 0; c < nc; ++c) v[c] = __ldcs(pred + s * M + m0 + c);
    upd(v, 1.0 / (double)(s - s0 + 1));
  }
  double* pm = part + (long long)g * 2 * M;
  for (int c = 0; c < nc; ++c) {
    pm[m0 + c] = mu[c];
    pm[M + m0 + c] = m2[c];
  }
}

static void moments_groups(long long S, long long M, int* G, int* Sg) {
  const long long quads = (M + 3) / 4;
  const long long blocks = (quads + kMomBlock - 1) / kMomBlock;
  const long long target = 4LL * sm_count();
  long long g = 1;
  if (blocks < target) g = (target + blocks - 1) / blocks;
  if (g > S) g = S;
  if (g > 65535) g = 65535;
  if (g < 1) g = 1;
  long long sg = (S + g - 1) / g;
  if (sg < 1) sg = 1;
  g = (S + sg - 1) / sg;
  if (g < 1) g = 1;
  *G = (int)g;
  *Sg = (int)sg;
}

}  // namespace bnn

using namespace bnn;

extern "C" int64_t bnn_moments_workspace_bytes(int64_t S, int64_t M) {
  int G, Sg;
  moments_groups(S, M, &G, &Sg);
  return (int64_t)G * 2 * M * (int64_t)sizeof(double);
}

extern "C" int32_t bnn_moments(const float* pred, int64_t S, int64_t M, float* mean, float* var, double* part_mean,
                               double* part_m2, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!pred || S < 1 || M < 1) BNN_FAIL("bnn_moments: bad arguments (S=%lld, M=%lld)", (long long)S, (long long)M);
  int G, Sg;
  moments_groups(S, M, &G, &Sg);
  const int64_t need = (int64_t)G * 2 * M * (int64_t)sizeof(double);
  if (!workspace || workspace_bytes < need) BNN_FAIL("bnn_moments: workspace needs %lld bytes", (long long)need);
  cudaStream_t st = (cudaStream_t)stream;
  const long long quads = (M + 3) / 4;
  dim3 grid((unsigned)((quads + kMomBlock - 1) / kMomBlock), (unsigned)G);
  moments_kernel<<<grid, kMomBlock, 0, st>>>(pred, S, M, Sg, (double*)workspace);
  BNN_CUDA(cudaGetLastError());
  return finalize_moments((const double*)workspace, nullptr, G, Sg, S, M, mean, var, part_mean, part_m2, st);
}

// R shards laid out as part_mean[R][M], part_m2[R][M] (e.g. the all-gathered per-GPU moments); the shard sample counts
// travel by value in the kernel arguments (no device buffer, graph-capturable)
struct CountsArg { long long c[64]; };
__global__ void merge_kernel(const double* __restrict__ pmean, const double* __restrict__ pm2, CountsArg counts, int R,
                                   long long M, float* __restrict__ mean, float* __restrict__ var,
                                   double* __restrict__ out_mean = nullptr, double* __restrict__ out_m2 = nullptr) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += stride) {
    double n = 0.0, mu = 0.0, m2 = 0.0;
    for (int r = 0; r < R; ++r) {
      const double nb = (double)counts.c[r];
      if (nb <= 0.0) continue;
      const double d = pmean[(long long)r * M + m] - mu;
      const double nn = n + nb;
      mu += d * (nb / nn);
      m2 += pm2[(long long)r * M + m] + d * d * (n * nb / nn);
      n = nn;
    }
    if (mean) mean[m] = (float)mu;
    if (var) var[m] = (float)(m2 / (n - 1.0));
    if (out_mean) out_mean[m] = mu;
    if (out_m2) out_m2[m] = m2;
  }
}

// same merge, but the result stays a float64 (mean, M2) pair -- one shard made of several chunks (host-buffer path)
extern "C" int32_t bnn_moments_merge_partials(const double* part_mean, const double* part_m2, const int64_t* counts_host, int32_t R,
                                              int64_t M, double* out_mean, double* out_m2, void* stream) {
  if (!part_mean || !part_m2 || !counts_host || R < 1 || R > 64 || M < 1 || (!out_mean && !out_m2))
    BNN_FAIL("bnn_moments_merge_partials: bad arguments (R=%d)", R);
  CountsArg c;
  for (int r = 0; r < 64; ++r) c.c[r] = r < R ? counts_host[r] : 0;
  merge_kernel<<<grid_for(M, 256), 256, 0, (cudaStream_t)stream>>>(part_mean, part_m2, c, R, M, nullptr, nullptr, out_mean, out_m2);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t bnn_moments_merge(const double* part_mean, const double* part_m2, const int64_t* counts_host, int32_t R,
                                     int64_t M, float* mean, float* var, void* stream) {
  if (!part_mean || !part_m2 || !counts_host || R < 1 || R > 64 || M < 1) BNN_FAIL("bnn_moments_merge: bad arguments (R=%d)", R);
  CountsArg c;
  for (int r = 0; r < 64; ++r) c.c[r] = r < R ? counts_host[r] : 0;
  merge_kernel<<<grid_for(M, 256), 256, 0, (cudaStream_t)stream>>>(part_mean, part_m2, c, R, M, mean, var);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

// ---- metrics -------------------------------------------------------------------
// acc[0..nout) = sum (mean - y)^2 ; acc[nout..2nout) = sum nll terms ; acc[2nout] = #(v <= 0)
namespace bnn {
constexpr int kMetBlock = 256;
constexpr int kMetMaxOut = 64;

__global__ void metrics_zero_kernel(double* acc, int n) {
  if (threadIdx.x < n) acc[threadIdx.x] = 0.0;
}

__global__ void __launch_bounds__(kMetBlock) metrics_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                                            const float* __restrict__ y, const float* __restrict__ noise_var,
                                                            long long N, int nout, double* __restrict__ acc) {
  __shared__ double sh[2 * kMetMaxOut + 1];
  for (int i = threadIdx.x; i < 2 * nout + 1; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const long long M = N * nout;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float half_log_2pi = 0.91893853320467274178f;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += stride) {
    const int o = (int)(m % nout);
    const float mu = mean[m], yy = y[m];
    float v = var[m];
    if (noise_var) v += noise_var[o];
    const float e = mu - yy;
    // -log N(y | mu, sqrt(v)) as torch evaluates it: (y-mu)^2 / (2 v) + log(sqrt(v)) + log(sqrt(2 pi))
    const float sd = sqrtf(v);
    const float nll = (yy - mu) * (yy - mu) / (2.0f * (sd * sd)) + logf(sd) + half_log_2pi;
    atomicAdd(&sh[o], (double)(e * e));
    atomicAdd(&sh[nout + o], (double)nll);
    if (!(v > 0.0f)) atomicAdd(&sh[2 * nout], 1.0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * nout + 1; i += blockDim.x) atomicAdd(&acc[i], sh[i]);
}

__global__ void metrics_final_kernel(const double* acc, long long N, int nout, float* out) {
  const int i = threadIdx.x;
  if (i < nout) {
    out[i] = (float)sqrt(acc[i] / (double)N);
    out[nout + i] = (float)(acc[nout + i] / (double)N);
  }
  if (i == 0) out[2 * nout] = (float)acc[2 * nout];
}
}  // namespace bnn

extern "C" int32_t bnn_metrics(const float* mean, const float* var, const float* y, const float* noise_var, int64_t N,
                               int32_t nout, float* out, double* scratch, void* stream) {
  if (!mean || !var || !y || !out || !scratch || N < 1 || nout < 1 || nout > kMetMaxOut)
    BNN_FAIL("bnn_metrics: bad arguments (N=%lld, nout=%d, max nout %d)", (long long)N, nout, kMetMaxOut);
  cudaStream_t st = (cudaStream_t)stream;
  // FP64 accumulators: caller-provided device scratch (no process-wide state: models on different GPUs / streams in one
  // process must not share accumulators)
  double* acc = scratch;
  metrics_zero_kernel<<<1, 2 * kMetMaxOut + 1, 0, st>>>(acc, 2 * nout + 1);
  metrics_kernel<<<grid_for(N * nout, kMetBlock * 4), kMetBlock, 0, st>>>(mean, var, y, noise_var, N, nout, acc);
  metrics_final_kernel<<<1, kMetMaxOut, 0, st>>>(acc, N, nout, out);
  BNN_CUDA(cudaGetLastError());
  return 0;
}
