// Shared host/device helpers for libbnn_b200: error plumbing, the packed
// parameter layout of include/bnn_b200.h, and sm_100a PTX wrappers
// (mbarrier + cp.async.bulk, i.e. the TMA bulk-copy engine).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bnn_b200.h"

namespace bnn {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
#define BNN_FAIL(...)                \
  do {                               \
    ::bnn::set_error(__VA_ARGS__);   \
    return -1;                       \
  } while (0)
#define BNN_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      ::bnn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return -2;                                                                    \
    }                                                                               \
  } while (0)

int sm_count();

// ---- packed layout --------------------------------------------------------
__host__ __device__ inline int roundup4(int v) { return (v + 3) & ~3; }

struct Layout {
  int n_lin;
  int in[BNN_MAX_LINEAR], out[BNN_MAX_LINEAR], ld[BNN_MAX_LINEAR];
  long long off[BNN_MAX_LINEAR + 1];  // off[n_lin] == stride
  int act;
  int wmax;  // widest of dims[]
};

// returns 0 on success
inline int make_layout(const BnnMlpDesc* d, Layout* L) {
  if (!d) { set_error("null BnnMlpDesc"); return -1; }
  if (d->n_linear < 1 || d->n_linear > BNN_MAX_LINEAR) {
    set_error("n_linear=%d out of range [1,%d]", d->n_linear, BNN_MAX_LINEAR);
    return -1;
  }
  if (d->act < BNN_ACT_RELU || d->act > BNN_ACT_IDENTITY) { set_error("unknown activation %d", d->act); return -1; }
  L->n_lin = d->n_linear;
  L->act = d->act;
  L->wmax = 0;
  long long o = 0;
  for (int l = 0; l <= d->n_linear; ++l) {
    if (d->dims[l] < 1 || d->dims[l] > BNN_MAX_WIDTH) {
      set_error("dims[%d]=%d out of range [1,%d]", l, d->dims[l], BNN_MAX_WIDTH);
      return -1;
    }
    if (d->dims[l] > L->wmax) L->wmax = d->dims[l];
  }
  for (int l = 0; l < d->n_linear; ++l) {
    L->in[l] = d->dims[l];
    L->out[l] = d->dims[l + 1];
    L->ld[l] = roundup4(d->dims[l] + 1);
    L->off[l] = o;
    o += (long long)L->out[l] * L->ld[l];
  }
  L->off[d->n_linear] = o;
  return 0;
}

// offsets of each masked layer's input mask inside the concatenated mask vector
inline int make_mask_offsets(const Layout& L, const int32_t* layer_masked, int* mask_off /*[BNN_MAX_LINEAR]*/) {
  int o = 0;
  for (int l = 0; l < BNN_MAX_LINEAR; ++l) {
    if (l < L.n_lin && layer_masked && layer_masked[l]) {
      mask_off[l] = o;
      o += L.in[l];
    } else {
      mask_off[l] = -1;
    }
  }
  return o;
}

// ---- activations ----------------------------------------------------------
__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case BNN_ACT_RELU: return fmaxf(v, 0.0f);
    case BNN_ACT_TANH: return tanhf(v);          // accurate libdevice tanh (<= 2 ulp), not tanh.approx
    case BNN_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default: return v;
  }
}

// ---- mbarrier / bulk copy (sm_90+ PTX, SASS: SYNCS.*, UBLKCP) ---------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// explicit shared-space vector accesses on 32-bit shared addresses (keeps ptxas from falling back to generic LD/ST)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// producer-side wait: the barrier completes only after a whole stage of FMAs, so back off between probes
// instead of burning issue slots of the SMSP that also hosts compute warps
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (done) break;
    __nanosleep(200);
  }
}
// Wait for the phase with the given parity.  try_wait takes a suspend-time hint: the hardware parks the thread until the
// phase completes or the time limit expires, instead of returning after a short default limit -- with a dozen waiting
// warps per SM, bare polling loops (TRYWAIT + BRA) took as many issue slots as the FFMAs of the working warps.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(1000000u)
      : "memory");
}
// latency-critical wait: plain polling (no suspend hint), for the few warps whose reaction time sits on the critical path
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITS_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONES_%=;\n"
      "bra WAITS_%=;\n"
      "DONES_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// non-blocking probe of a phase (true = complete); lets a single-warp pipeline overlap the probe latency with other work
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, both 16B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace bnn
