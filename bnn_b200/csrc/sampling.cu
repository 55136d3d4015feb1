// Kernels (b), (c), (d): streaming, HBM-bound elementwise work over the packed
// parameter layout.  One thread = 4 consecutive floats (one 16-byte access per
// operand, one Philox block -> 4 normals).
//   (b) bnn_reparam_kl   : W[s] = mu + sigma(rho) * eps   + KL(q || N(0, weight_std))
//                          BNN_BBB.py:24-27, :107-110; BNN_SVI.py:42-46
//   (c) bnn_psgld_step   : preconditioned SGLD update (pybnn's pSGLD.step at
//                          BNN_SGDMC.py:86; arithmetic restated from Li et al. 2016) on a caller-supplied
//                          gradient; the chain itself runs the fused step of sgld_step.cu, whose update
//                          epilogue is the same arithmetic
//   (d) bnn_dropout_masks / bnn_apply_masks : BNN_Dropout.py:70-75, BNN_CDropout.py:59-65
#include "common.cuh"
#include "philox.cuh"

namespace bnn {

struct LayoutDev {
  int n_lin;
  int in[BNN_MAX_LINEAR], out[BNN_MAX_LINEAR], ld[BNN_MAX_LINEAR];
  long long off[BNN_MAX_LINEAR + 1];
  int mask_off[BNN_MAX_LINEAR];
  float p_drop[BNN_MAX_LINEAR];
};

static LayoutDev to_dev(const Layout& L) {
  LayoutDev d;
  memset(&d, 0, sizeof(d));
  d.n_lin = L.n_lin;
  for (int l = 0; l < L.n_lin; ++l) { d.in[l] = L.in[l]; d.out[l] = L.out[l]; d.ld[l] = L.ld[l]; d.off[l] = L.off[l]; }
  d.off[L.n_lin] = L.off[L.n_lin];
  for (int l = 0; l < BNN_MAX_LINEAR; ++l) d.mask_off[l] = -1;
  return d;
}

// layer and first column of the 4-float group starting at flat element e (groups never straddle rows)
__device__ __forceinline__ void locate(const LayoutDev& L, long long e, int& layer, int& col) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < BNN_MAX_LINEAR; ++i)
    if (i < L.n_lin && e >= L.off[i]) l = i;
  layer = l;
  col = (int)((e - L.off[l]) % L.ld[l]);
}

__device__ __forceinline__ float softplus_t20(float x) { return x > 20.0f ? x : log1pf(expf(x)); }  // F.softplus default

__device__ __forceinline__ float scale_of(float rho, int mode) {
  return mode == BNN_SCALE_BBB ? sqrtf(softplus_t20(fmaxf(rho, -35.0f))) : softplus_t20(rho);
}

// ---- (b) -------------------------------------------------------------------------
constexpr int kEwBlock = 256;

__global__ void __launch_bounds__(kEwBlock) reparam_kernel(const __grid_constant__ LayoutDev L, int mode,
                                                           const float* __restrict__ mu, const float* __restrict__ rho,
                                                           const float* __restrict__ eps, unsigned long long seed,
                                                           long long sample_offset, long long S, float* __restrict__ W,
                                                           float inv_wstd, double* __restrict__ kl_out) {
  const long long P = L.off[L.n_lin];
  const long long quads = P >> 2;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double kl_local = 0.0;
  if (q < quads) {
    const long long e = q << 2;
    int layer, col;
    locate(L, e, layer, col);
    const int ncols = L.in[layer] + 1;
    const float4 m4 = *reinterpret_cast<const float4*>(mu + e);
    const float4 r4 = *reinterpret_cast<const float4*>(rho + e);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w};
    const float r[4] = {r4.x, r4.y, r4.z, r4.w};
    float sg[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) sg[c] = scale_of(r[c], mode);
    // samples handled by this thread: blockIdx.y, blockIdx.y + gridDim.y, ...
    // two samples per trip: two independent Philox / Box-Muller chains in flight per thread (the kernel is bound by the
    // ~200 instructions per quad of the counter RNG and the accurate logf / sincospif, not by its 16-byte store)
#pragma unroll 2
    for (long long s = blockIdx.y; s < S; s += gridDim.y) {
      float z[4];
      if (eps) {
        const float4 e4 = __ldcs(reinterpret_cast<const float4*>(eps + s * P + e));
        z[0] = e4.x; z[1] = e4.y; z[2] = e4.z; z[3] = e4.w;
      } else {
        philox_normal4(seed, (uint32_t)(sample_offset + s), STREAM_REPARAM, (uint64_t)q, z);
      }
      float w[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) w[c] = (col + c < ncols) ? __fadd_rn(m[c], __fmul_rn(sg[c], z[c])) : 0.0f;  // mul then add, as torch does
      __stcs(reinterpret_cast<float4*>(W + s * P + e), make_float4(w[0], w[1], w[2], w[3]));
    }
    if (kl_out && blockIdx.y == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (col + c < ncols) {
          // torch _kl_normal_normal: 0.5 * (var_ratio + t1 - 1 - log(var_ratio))
          const float sr = sg[c] * inv_wstd;
          const float var_ratio = sr * sr;
          const float tm = m[c] * inv_wstd;
          const float t1 = tm * tm;
          kl_local += (double)(0.5f * (var_ratio + t1 - 1.0f - logf(var_ratio)));
        }
      }
    }
  }
  if (kl_out && blockIdx.y == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kl_local += __shfl_xor_sync(0xffffffffu, kl_local, o);
    __shared__ double wsum[kEwBlock / 32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = kl_local;
    __syncthreads();
    if (threadIdx.x < 32) {
      double v = threadIdx.x < kEwBlock / 32 ? wsum[threadIdx.x] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0) atomicAdd(kl_out, v);
    }
  }
}

// ---- (d) -------------------------------------------------------------------------
// thread = one Philox block = 4 consecutive entries of one sample's concatenated mask vector (element j lives in word j & 3
// of block j >> 2, as in the in-kernel paths of the forwards); samples along blockIdx.y.  (One thread per ENTRY recomputed the
// block four times and paid a 64-bit division per entry: 620 GB/s at S = 200k; ncu: ALU pipe 63 %.)
__global__ void __launch_bounds__(kEwBlock) masks_kernel(const __grid_constant__ LayoutDev L, int mode, int mask_len,
                                                         const float* __restrict__ uniforms, float temperature,
                                                         unsigned long long seed, long long sample_offset, long long S,
                                                         float* __restrict__ out) {
  const int qd = blockIdx.x * blockDim.x + threadIdx.x;
  const int j0 = qd << 2;
  if (j0 >= mask_len) return;
  const uint32_t tag = mode == BNN_MASK_BERNOULLI ? STREAM_BERNOULLI : STREAM_CONCRETE;
  // layer (-> drop probability) of each of the four entries: the same for every sample
  float pd[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int layer = 0;
#pragma unroll
    for (int l = 0; l < BNN_MAX_LINEAR; ++l)
      if (L.mask_off[l] >= 0 && j0 + e >= L.mask_off[l]) layer = l;
    pd[e] = L.p_drop[layer];
  }
  for (long long s = blockIdx.y; s < S; s += gridDim.y) {
    u32x4 blk = u32x4{0u, 0u, 0u, 0u};
    if (!uniforms) blk = philox_block(seed, (uint32_t)(sample_offset + s), tag, (uint64_t)qd);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j0 + e;
      if (j < mask_len) {
        const float u = uniforms ? uniforms[s * mask_len + j] : u24(pick_word(blk, e));
        out[s * mask_len + j] = mode == BNN_MASK_BERNOULLI ? bernoulli_keep(u, pd[e]) : concrete_keep(u, pd[e], temperature);
      }
    }
  }
}

__global__ void __launch_bounds__(kEwBlock) apply_masks_kernel(const __grid_constant__ LayoutDev L, const float* __restrict__ base,
                                                               const float* __restrict__ masks, long long mask_stride,
                                                               long long S, float* __restrict__ W) {
  const long long P = L.off[L.n_lin];
  const long long quads = P >> 2;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= quads) return;
  const long long e = q << 2;
  int layer, col;
  locate(L, e, layer, col);
  const float4 b4 = *reinterpret_cast<const float4*>(base + e);
  const float b[4] = {b4.x, b4.y, b4.z, b4.w};
  const int mo = L.mask_off[layer];
  const int nin = L.in[layer];
  for (long long s = blockIdx.y; s < S; s += gridDim.y) {
    float w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int cc = col + c;
      w[c] = (mo >= 0 && cc < nin) ? b[c] * __ldg(masks + s * mask_stride + mo + cc) : b[c];
    }
    __stcs(reinterpret_cast<float4*>(W + s * P + e), make_float4(w[0], w[1], w[2], w[3]));
  }
}

// ---- (c) -------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwBlock) psgld_kernel(float* __restrict__ theta, const float* __restrict__ grad,
                                                         float* __restrict__ V, long long n, long long n_head, float lr_head,
                                                         float lr_tail, float alpha, float lam, const float* __restrict__ noise,
                                                         unsigned long long seed, long long step) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long e = q << 2;
  if (e >= n) return;
  float z[4];
  const float om = 1.0f - alpha;
  if (e + 4 <= n) {
    // the three streaming loads go out first; the ~200 instructions of Philox + Box-Muller run under their latency
    const float4 t4 = __ldcs(reinterpret_cast<const float4*>(theta + e));
    const float4 g4 = __ldcs(reinterpret_cast<const float4*>(grad + e));
    const float4 v4 = __ldcs(reinterpret_cast<const float4*>(V + e));
    if (!noise) philox_normal4(seed, (uint32_t)step, STREAM_SGLD, (uint64_t)q, z);
    if (noise) {
      const float4 n4 = *reinterpret_cast<const float4*>(noise + e);
      z[0] = n4.x; z[1] = n4.y; z[2] = n4.z; z[3] = n4.w;
    }
    float t[4] = {t4.x, t4.y, t4.z, t4.w};
    const float g[4] = {g4.x, g4.y, g4.z, g4.w};
    float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float lr = (e + c) < n_head ? lr_head : lr_tail;
      v[c] = alpha * v[c] + om * g[c] * g[c];
      const float G = 1.0f / (lam + sqrtf(v[c]));
      t[c] = t[c] - (0.5f * lr * G * g[c] + sqrtf(lr * G) * z[c]);
    }
    __stcs(reinterpret_cast<float4*>(theta + e), make_float4(t[0], t[1], t[2], t[3]));
    __stcs(reinterpret_cast<float4*>(V + e), make_float4(v[0], v[1], v[2], v[3]));
  } else {
    if (!noise) philox_normal4(seed, (uint32_t)step, STREAM_SGLD, (uint64_t)q, z);
    for (int c = 0; e + c < n; ++c) {
      const float lr = (e + c) < n_head ? lr_head : lr_tail;
      const float g = grad[e + c];
      const float v = alpha * V[e + c] + om * g * g;
      const float G = 1.0f / (lam + sqrtf(v));
      const float zz = noise ? noise[e + c] : z[c];
      theta[e + c] = theta[e + c] - (0.5f * lr * G * g + sqrtf(lr * G) * zz);
      V[e + c] = v;
    }
  }
}

// ---- raw stream dumps (tests) --------------------------------------------------------
__global__ void philox_words_kernel(unsigned long long seed, uint32_t subseq, uint32_t tag, long long n, uint32_t* out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((q << 2) >= n) return;
  const u32x4 v = philox_block(seed, subseq, tag, (uint64_t)q);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  for (int c = 0; c < 4 && (q << 2) + c < n; ++c) out[(q << 2) + c] = w[c];
}
__global__ void philox_normals_kernel(unsigned long long seed, uint32_t subseq, uint32_t tag, long long n, float* out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((q << 2) >= n) return;
  float z[4];
  philox_normal4(seed, subseq, tag, (uint64_t)q, z);
  for (int c = 0; c < 4 && (q << 2) + c < n; ++c) out[(q << 2) + c] = z[c];
}

static unsigned sample_rows(long long S, long long blocks_x) {
  // y-dimension: ~32 CTAs per SM in total (several CTAs are resident per SM, so this is ~6-8 waves and the ragged last
  // wave costs little -- 4 per SM came out as 1.04 waves on the cfg 4 bank and doubled the run time), each CTA looping
  // over its samples
  long long want = (32LL * sm_count() + blocks_x - 1) / blocks_x;
  if (want < 1) want = 1;
  if (want > S) want = S;
  if (want > 65535) want = 65535;
  return (unsigned)want;
}

}  // namespace bnn

using namespace bnn;

extern "C" int32_t bnn_reparam_kl(const BnnMlpDesc* d, int32_t scale_mode, const float* mu, const float* rho, const float* eps,
                                  uint64_t seed, int64_t sample_offset, int64_t S, float* W_out, float weight_std, double* kl_out,
                                  void* stream) {
  Layout L;
  if (make_layout(d, &L)) return -1;
  if (!mu || !rho || (S > 0 && !W_out) || S < 0) BNN_FAIL("bnn_reparam_kl: bad arguments");
  if (scale_mode != BNN_SCALE_BBB && scale_mode != BNN_SCALE_SVI) BNN_FAIL("bnn_reparam_kl: unknown scale mode %d", scale_mode);
  if (kl_out && !(weight_std > 0.f)) BNN_FAIL("bnn_reparam_kl: weight_std must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (kl_out) BNN_CUDA(cudaMemsetAsync(kl_out, 0, sizeof(double), st));
  const long long quads = L.off[L.n_lin] >> 2;
  const long long bx = (quads + kEwBlock - 1) / kEwBlock;
  if (S == 0 && !kl_out) return 0;
  dim3 grid((unsigned)bx, S > 0 ? sample_rows(S, bx) : 1);
  reparam_kernel<<<grid, kEwBlock, 0, st>>>(to_dev(L), scale_mode, mu, rho, eps, seed, sample_offset, S, W_out,
                                            kl_out ? 1.0f / weight_std : 0.f, kl_out);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t bnn_dropout_masks(const BnnMlpDesc* d, const BnnMaskSpec* spec, const float* uniforms, int64_t S,
                                     float* masks_out, void* stream) {
  Layout L;
  if (make_layout(d, &L)) return -1;
  if (!spec || !masks_out || S < 1) BNN_FAIL("bnn_dropout_masks: bad arguments");
  if (spec->mode != BNN_MASK_BERNOULLI && spec->mode != BNN_MASK_CONCRETE)
    BNN_FAIL("bnn_dropout_masks: mode must be BNN_MASK_BERNOULLI or BNN_MASK_CONCRETE");
  LayoutDev ld = to_dev(L);
  const int mask_len = make_mask_offsets(L, spec->layer_masked, ld.mask_off);
  if (mask_len < 1) BNN_FAIL("bnn_dropout_masks: no layer is masked");
  for (int l = 0; l < L.n_lin; ++l) ld.p_drop[l] = spec->p_drop[l];
  const long long bx = ((mask_len + 3) / 4 + kEwBlock - 1) / kEwBlock;
  dim3 grid((unsigned)bx, sample_rows(S, bx));
  masks_kernel<<<grid, kEwBlock, 0, (cudaStream_t)stream>>>(ld, spec->mode, mask_len, uniforms, spec->temperature,
                                                            spec->seed, spec->sample_offset, S, masks_out);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t bnn_apply_masks(const BnnMlpDesc* d, const int32_t* layer_masked, const float* base, const float* masks,
                                   int64_t mask_stride, int64_t S, float* W_out, void* stream) {
  Layout L;
  if (make_layout(d, &L)) return -1;
  if (!layer_masked || !base || !masks || !W_out || S < 1) BNN_FAIL("bnn_apply_masks: bad arguments");
  LayoutDev ld = to_dev(L);
  make_mask_offsets(L, layer_masked, ld.mask_off);
  const long long quads = L.off[L.n_lin] >> 2;
  const long long bx = (quads + kEwBlock - 1) / kEwBlock;
  dim3 grid((unsigned)bx, sample_rows(S, bx));
  apply_masks_kernel<<<grid, kEwBlock, 0, (cudaStream_t)stream>>>(ld, base, masks, mask_stride, S, W_out);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t bnn_psgld_step(float* theta, const float* grad, float* V, int64_t n, int64_t n_head, float lr_head,
                                  float lr_tail, float alpha, float lam, const float* noise, uint64_t seed, int64_t step,
                                  void* stream) {
  if (!theta || !grad || !V || n < 1) BNN_FAIL("bnn_psgld_step: bad arguments");
  if ((((uintptr_t)theta | (uintptr_t)grad | (uintptr_t)V | (uintptr_t)noise) & 15) != 0)
    BNN_FAIL("bnn_psgld_step: pointers must be 16-byte aligned");
  const long long quads = (n + 3) >> 2;
  psgld_kernel<<<(unsigned)((quads + kEwBlock - 1) / kEwBlock), kEwBlock, 0, (cudaStream_t)stream>>>(
      theta, grad, V, n, n_head, lr_head, lr_tail, alpha, lam, noise, seed, step);
  BNN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int32_t bnn_philox_words(uint64_t seed, uint32_t subsequence, uint32_t tag, int64_t n, uint32_t* out, void* stream) {
  if (!out || n < 1) BNN_FAIL("bnn_philox_words: bad arguments");
  const long long quads = (n + 3) >> 2;
  philox_words_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, subsequence, tag, n, out);
  BNN_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int32_t bnn_philox_normals(uint64_t seed, uint32_t subsequence, uint32_t tag, int64_t n, float* out, void* stream) {
  if (!out || n < 1) BNN_FAIL("bnn_philox_normals: bad arguments");
  const long long quads = (n + 3) >> 2;
  philox_normals_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, subsequence, tag, n, out);
  BNN_CUDA(cudaGetLastError());
  return 0;
}
