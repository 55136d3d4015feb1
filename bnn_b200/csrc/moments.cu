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

// Few points, many sample groups (cfg 1: 51 points x 100 groups; the serial loop above then is 100 dependent FP64 merges with
// two divisions each on 51 threads -- 60 us for a 12 us forward): ONE WARP per point, lane = a slice of the groups, partial
// Chan merges per lane, then a five-level butterfly.  Lane 0 writes.
__device__ __forceinline__ void chan_merge(double& n, double& mu, double& m2, double nb, double mb, double m2b) {
  const double nn = n + nb;
  if (nn <= 0.0) return;
  const double inv = 1.0 / nn, d = mb - mu;
  mu += d * (nb * inv);
  m2 += m2b + d * d * (n * nb * inv);
  n = nn;
}
__global__ void __launch_bounds__(256) finalize_wide_kernel(const double* __restrict__ part, int G, int Sg, long long S, long long M,
                                                            float* __restrict__ mean, float* __restrict__ var,
                                                            double* __restrict__ out_mean, double* __restrict__ out_m2) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M; m += warps) {
    double n = 0.0, mu = 0.0, m2 = 0.0;
    for (int g = lane; g < G; g += 32) {
      const long long rem = S - (long long)g * Sg;
      const double nb = (double)(rem < Sg ? rem : Sg);
      if (nb > 0.0) chan_merge(n, mu, m2, nb, part[(long long)g * 2 * M + m], part[(long long)g * 2 * M + M + m]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mu, o),
                   m2b = __shfl_xor_sync(0xffffffffu, m2, o);
      chan_merge(n, mu, m2, nb, mb, m2b);
    }
    if (lane == 0) {
      if (mean) mean[m] = (float)mu;
      if (var) var[m] = (float)(m2 / (n - 1.0));  // n == 1 -> 0/0 = NaN, as torch.var
      if (out_mean) out_mean[m] = mu;
      if (out_m2) out_m2[m] = m2;
    }
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
  if (G >= 8 && M <= 32768)   // launch-bound sizes: a warp per point
    finalize_wide_kernel<<<grid_for(M * 32, 256), 256, 0, st>>>(part, G, Sg, S, M, mean, var, part_mean, part_m2);
  else
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

// ---- multi-GPU exchange of the per-point moments over NCCL ---------------------------------------------
// Each rank holds float64 (mean, M2) of ITS samples for all M points.  Chan's merge is not a plain sum, but float64 raw moments
// are: sum_x = n mean, sum_xx = M2 + n mean^2.  Formed from float64 Welford state and reduced in float64 they lose nothing that
// matters (var = (sum_xx - sum_x^2 / n) / (n - 1) cancels ~mean^2 / var digits of 1e-16: 2.5e-11 relative on the mean 5 /
// sigma 0.01 stress case; float32 raw moments lose 13 % there).  So ONE ncclReduceScatter(ncclSum) of [R][2 Mslice + 1] doubles
// leaves every rank with the totals of its slice of points (+ the total sample count in the last element), a kernel finalises the
// slice, and an optional ncclAllGather of 8 bytes per point hands everyone the full mean / var.  The payload per rank is
// 16 M bytes whatever the number of ranks (the all-gather-then-merge form lands 16 M R bytes on every rank).  No host sync.
// NCCL is reached through dlopen (the library has no link-time dependency on it): the communicator must come from the same
// libnccl, see bnn_nccl_load.
namespace bnn {
typedef int (*NcclReduceScatterFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*NcclAllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*NcclCommCountFn)(void*, int*);
typedef int (*NcclCommUserRankFn)(void*, int*);
static void* g_nccl = nullptr;
static NcclReduceScatterFn p_reduce_scatter = nullptr;
static NcclAllGatherFn p_all_gather = nullptr;
static NcclCommCountFn p_comm_count = nullptr;
static NcclCommUserRankFn p_comm_rank = nullptr;
constexpr int kNcclFloat = 7, kNcclDouble = 8, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)

// sendbuf[r][0][i] = n mean, [r][1][i] = M2 + n mean^2 for point r * Ms + i (zero beyond M); sendbuf[r][2 Ms] = n
__global__ void nccl_pack_kernel(const double* __restrict__ pm, const double* __restrict__ pm2, double n, long long M, long long Ms,
                                 int R, double* __restrict__ send) {
  const long long per = 2 * Ms + 1;
  const long long total = (long long)R * Ms;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Ms, j = i - r * Ms;
    double sx = 0.0, sxx = 0.0;
    if (i < M) { const double mu = pm[i]; sx = n * mu; sxx = pm2[i] + n * mu * mu; }
    send[r * per + j] = sx;
    send[r * per + Ms + j] = sxx;
    if (j == 0) send[r * per + 2 * Ms] = n;
  }
}
// this rank's slice: totals -> (mean, unbiased var), float32, laid out [2][Ms] for the all-gather
__global__ void nccl_finalize_kernel(const double* __restrict__ recv, long long Ms, float* __restrict__ out) {
  const double n = recv[2 * Ms];
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < Ms; j += (long long)gridDim.x * blockDim.x) {
    const double sx = recv[j], sxx = recv[Ms + j];
    const double mu = sx / n;
    out[j] = (float)mu;
    out[Ms + j] = (float)((sxx - sx * mu) / (n - 1.0));     // n == 1 -> 0 / 0 = NaN, as torch.var
  }
}
__global__ void nccl_unpack_kernel(const float* __restrict__ g, long long M, long long Ms, float* __restrict__ mean, float* __restrict__ var) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Ms, j = i - r * Ms;
    if (mean) mean[i] = g[r * 2 * Ms + j];
    if (var) var[i] = g[r * 2 * Ms + Ms + j];
  }
}
}  // namespace bnn

#include <dlfcn.h>

// dlopen the NCCL the caller's communicator comes from (path may be NULL: "libnccl.so.2" on the default search path)
extern "C" int32_t bnn_nccl_load(const char* path) {
  if (g_nccl) return 0;
  void* h = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) BNN_FAIL("bnn_nccl_load: %s", dlerror());
  p_reduce_scatter = (NcclReduceScatterFn)dlsym(h, "ncclReduceScatter");
  p_all_gather = (NcclAllGatherFn)dlsym(h, "ncclAllGather");
  p_comm_count = (NcclCommCountFn)dlsym(h, "ncclCommCount");
  p_comm_rank = (NcclCommUserRankFn)dlsym(h, "ncclCommUserRank");
  if (!p_reduce_scatter || !p_all_gather || !p_comm_count || !p_comm_rank) BNN_FAIL("bnn_nccl_load: NCCL symbols missing");
  g_nccl = h;
  return 0;
}

extern "C" int64_t bnn_moments_allreduce_workspace_bytes(int64_t M, int32_t n_ranks) {
  if (M < 1 || n_ranks < 1) { set_error("bnn_moments_allreduce_workspace_bytes: bad arguments"); return -1; }
  const long long Ms = (M + n_ranks - 1) / n_ranks;
  // send [R][2 Ms + 1] doubles | recv [2 Ms + 1] doubles | slice [2 Ms] floats | gathered [R][2 Ms] floats
  return (long long)n_ranks * (2 * Ms + 1) * 8 + (2 * Ms + 1) * 8 + 2 * Ms * 4 + (long long)n_ranks * 2 * Ms * 4 + 64;
}

extern "C" int32_t bnn_moments_allreduce(void* nccl_comm, const double* part_mean, const double* part_m2, int64_t count, int64_t M,
                                         float* mean, float* var, int32_t gather, void* workspace, int64_t workspace_bytes,
                                         void* stream) {
  if (!nccl_comm || !part_mean || !part_m2 || count < 1 || M < 1 || !workspace) BNN_FAIL("bnn_moments_allreduce: bad arguments");
  if (!g_nccl && bnn_nccl_load(nullptr)) return -1;
  int R = 0, rank = 0;
  if (p_comm_count(nccl_comm, &R) != 0 || p_comm_rank(nccl_comm, &rank) != 0 || R < 1) BNN_FAIL("bnn_moments_allreduce: bad communicator");
  const long long Ms = (M + R - 1) / R;
  if (workspace_bytes < bnn_moments_allreduce_workspace_bytes(M, R)) BNN_FAIL("bnn_moments_allreduce: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* send = (double*)workspace;
  double* recv = send + (long long)R * (2 * Ms + 1);
  float* slice = (float*)(recv + (2 * Ms + 1));
  float* gathered = slice + 2 * Ms;
  nccl_pack_kernel<<<grid_for((long long)R * Ms, 256), 256, 0, st>>>(part_mean, part_m2, (double)count, M, Ms, R, send);
  BNN_CUDA(cudaGetLastError());
  if (p_reduce_scatter(send, recv, (size_t)(2 * Ms + 1), kNcclDouble, kNcclSum, nccl_comm, st) != 0) BNN_FAIL("ncclReduceScatter failed");
  nccl_finalize_kernel<<<grid_for(Ms, 256), 256, 0, st>>>(recv, Ms, slice);
  BNN_CUDA(cudaGetLastError());
  if (gather) {
    // every rank ends with the full mean / var (what BNN.predict_mv returns)
    if (p_all_gather(slice, gathered, (size_t)(2 * Ms), kNcclFloat, nccl_comm, st) != 0) BNN_FAIL("ncclAllGather failed");
    nccl_unpack_kernel<<<grid_for(M, 256), 256, 0, st>>>(gathered, M, Ms, mean, var);
    BNN_CUDA(cudaGetLastError());
  } else {
    // each rank keeps only the points [rank * Ms, min(M, (rank + 1) * Ms)) of mean / var (written at those positions)
    const long long lo = (long long)rank * Ms;
    const long long n = lo >= M ? 0 : (M - lo < Ms ? M - lo : Ms);
    if (n > 0 && mean) BNN_CUDA(cudaMemcpyAsync(mean + lo, slice, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    if (n > 0 && var) BNN_CUDA(cudaMemcpyAsync(var + lo, slice + Ms, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  }
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
  if (nout == 1) {
    // single output (every configuration but the multi-output driver): per-thread float64 sums, warp shuffles, one shared
    // atomic per warp (one shared-memory double atomic per POINT made this kernel 96 GB/s at N = 1M)
    double se = 0.0, sn = 0.0, bad = 0.0;
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += stride) {
      const float mu = mean[m], yy = y[m];
      float v = var[m];
      if (noise_var) v += noise_var[0];
      const float e = mu - yy;
      const float sd = sqrtf(v);
      se += (double)(e * e);
      sn += (double)((yy - mu) * (yy - mu) / (2.0f * (sd * sd)) + logf(sd) + half_log_2pi);
      if (!(v > 0.0f)) bad += 1.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      se += __shfl_xor_sync(0xffffffffu, se, o);
      sn += __shfl_xor_sync(0xffffffffu, sn, o);
      bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sh[0], se); atomicAdd(&sh[1], sn); atomicAdd(&sh[2], bad); }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(&acc[threadIdx.x], sh[threadIdx.x]);
    return;
  }
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
