// Counter-based RNG of the device path: Philox4x32-10 (Salmon et al. SC'11).
// The reference draws from torch's global mt19937 (BNN_BBB.py:25, BNN_Dropout.py:74,
// torch.rand inside RelaxedBernoulli for BNN_CDropout.py:64), which cannot be
// reproduced on a device; here every draw is a pure function of
// (seed, subsequence, stream tag, element index) so that results do not depend on
// the launch geometry or on how samples are sharded across GPUs.
// CPU twin: oracle/philox.py (bit-exact words / uniforms / Bernoulli masks).
#pragma once
#include <stdint.h>

namespace bnn {

enum : uint32_t { STREAM_REPARAM = 1, STREAM_BERNOULLI = 2, STREAM_CONCRETE = 3, STREAM_SGLD = 4 };

struct u32x4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return u32x4{c0, c1, c2, c3};
}

// block `blk` of stream (seed, subsequence, tag)
__host__ __device__ __forceinline__ u32x4 philox_block(uint64_t seed, uint32_t subseq, uint32_t tag, uint64_t blk) {
  return philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), subseq, tag, (uint32_t)seed, (uint32_t)(seed >> 32));
}

__host__ __device__ __forceinline__ uint32_t pick_word(const u32x4& v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// [0,1) with 24 random bits; exact in binary32
__host__ __device__ __forceinline__ float u24(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-08f; }

// Box-Muller on one word pair -> two standard normals
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, float& z0, float& z1) {
  const float u1 = (float)((wa >> 8) + 1u) * 5.9604644775390625e-08f;  // (0,1]
  const float u2 = (float)(wb >> 8) * 5.9604644775390625e-08f;         // [0,1)
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// four normals from one Philox block
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t subseq, uint32_t tag, uint64_t blk, float z[4]) {
  const u32x4 v = philox_block(seed, subseq, tag, blk);
  box_muller(v.x, v.y, z[0], z[1]);
  box_muller(v.z, v.w, z[2], z[3]);
}

// ---- dropout keep masks ------------------------------------------------------
// 0/1 keep mask of BNN_Dropout.py:74 (F.dropout(ones, p) * (1 - p)): keep iff u < 1 - p
__host__ __device__ __forceinline__ float bernoulli_keep(float u, float p_drop) { return u < (1.0f - p_drop) ? 1.0f : 0.0f; }

// relaxed-Bernoulli keep mask of BNN_CDropout.py:59-65 given the layer's drop probability
// p = clamp(sigmoid(p_logit)) (BNN_CDropout.py:18-19); follows torch's
// LogitRelaxedBernoulli.rsample + _clipped_sigmoid.
__device__ __forceinline__ float concrete_keep(float u, float p_drop, float temperature) {
  const float eps = 1.1920928955078125e-07f;   // finfo(float32).eps
  const float tiny = 1.1754943508222875e-38f;  // finfo(float32).tiny
  const float p = fminf(fmaxf(p_drop, eps), 1.0f - eps);
  const float q = fminf(fmaxf(1.0f - p, eps), 1.0f - eps);
  u = fminf(fmaxf(u, eps), 1.0f - eps);
  const float logit = (logf(u) - log1pf(-u) + logf(q) - log1pf(-q)) / temperature;
  const float z = 1.0f / (1.0f + expf(-logit));
  return fminf(fmaxf(z, tiny), 1.0f - eps);
}

}  // namespace bnn
