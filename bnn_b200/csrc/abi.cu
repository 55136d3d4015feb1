// C-ABI glue: error state, layout queries, forward dispatch and the host-buffer
// (end-to-end) convenience entry point.  No torch types anywhere.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace bnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  // cached per device: conf['device'] lets models sit on different GPUs of one process
  static int n[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;  // B200
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  }
  return n[dev];
}

int forward_simt(const BnnForwardArgs* a, cudaStream_t st);
long long forward_simt_workspace_bytes(const BnnForwardArgs* a);
int forward_tc(const BnnForwardArgs* a, cudaStream_t st);
long long forward_tc_workspace_bytes(const BnnForwardArgs* a);
int forward_tc_supported(const BnnForwardArgs* a);
int forward_tcl(const BnnForwardArgs* a, cudaStream_t st);
long long forward_tcl_workspace_bytes(const BnnForwardArgs* a);
int forward_tcl_supported(const BnnForwardArgs* a);

static bool use_tc(const BnnForwardArgs* a) {
  return a->precision == BNN_PREC_TF32X3 || a->precision == BNN_PREC_TF32 || a->precision == BNN_PREC_TF32_BF16;
}

// tensor-core requests: the resident-weights kernel (forward_tc.cu) when the shape fits it, else the streaming kernel for
// deep / wide networks (forward_tcl.cu); neither -> error with the first kernel's reason (no silent fallback)
static bool prefer_stream() {   // tuning knob: BNN_TC_STREAM=1 sends every supported shape to the streaming kernel
  const char* e = getenv("BNN_TC_STREAM");
  return e && atoi(e) != 0;
}
static long long tc_workspace_bytes(const BnnForwardArgs* a) {
  if (prefer_stream() && forward_tcl_supported(a)) return forward_tcl_workspace_bytes(a);
  if (forward_tc_supported(a)) return forward_tc_workspace_bytes(a);
  if (forward_tcl_supported(a)) return forward_tcl_workspace_bytes(a);
  forward_tc_supported(a);   // restore the first kernel's message
  return -1;
}
static int tc_forward(const BnnForwardArgs* a, cudaStream_t st) {
  if (prefer_stream() && forward_tcl_supported(a)) return forward_tcl(a, st);
  if (forward_tc_supported(a)) return forward_tc(a, st);
  if (forward_tcl_supported(a)) return forward_tcl(a, st);
  return forward_tc(a, st);  // fails with the reason
}

static int check_forward_args(const BnnForwardArgs* a, bool host) {
  if (!a) BNN_FAIL("bnn_forward: null args");
  if (a->N < 1 || a->S < 1) BNN_FAIL("bnn_forward: need N >= 1 and S >= 1 (N=%lld, S=%lld)", (long long)a->N, (long long)a->S);
  if (a->S > 0x7fffffffLL) BNN_FAIL("bnn_forward: S too large");
  if (!a->X || !a->W) BNN_FAIL("bnn_forward: X and W must not be null");
  Layout L;
  if (make_layout(&a->desc, &L)) return -1;
  if (a->w_stride != 0 && a->w_stride < L.off[L.n_lin])
    BNN_FAIL("bnn_forward: w_stride=%lld smaller than the packed network (%lld floats)", (long long)a->w_stride, L.off[L.n_lin]);
  if (a->w_stride & 3) BNN_FAIL("bnn_forward: w_stride must be a multiple of 4 floats");
  if (!host && (((uintptr_t)a->W) & 15)) BNN_FAIL("bnn_forward: W must be 16-byte aligned");
  if (a->mask.mode < BNN_MASK_NONE || a->mask.mode > BNN_MASK_CONCRETE) BNN_FAIL("bnn_forward: bad mask mode %d", a->mask.mode);
  if (a->mask.mode == BNN_MASK_EXTERNAL && !a->mask.mask) BNN_FAIL("bnn_forward: BNN_MASK_EXTERNAL without a mask tensor");
  if (a->precision < BNN_PREC_FP32 || a->precision > BNN_PREC_TF32_BF16) BNN_FAIL("bnn_forward: unknown precision %d", a->precision);
  if (!a->pred && !a->mean && !a->var && !a->part_mean && !a->part_m2) BNN_FAIL("bnn_forward: no output requested");
  return 0;
}

}  // namespace bnn

using namespace bnn;

extern "C" int32_t bnn_abi_version(void) { return BNN_ABI_VERSION; }
extern "C" const char* bnn_last_error(void) { return g_err; }
extern "C" int32_t bnn_sm_count(void) { return sm_count(); }
extern "C" int32_t bnn_forward_tc_supported(const BnnForwardArgs* a) {
  if (!a) { set_error("null args"); return 0; }
  if (forward_tc_supported(a)) return 1;
  if (forward_tcl_supported(a)) return 1;
  forward_tc_supported(a);   // leave the resident-weights kernel's reason in bnn_last_error()
  return 0;
}
extern "C" int32_t bnn_set_device(int32_t device) {
  BNN_CUDA(cudaSetDevice(device));
  return 0;
}

extern "C" int64_t bnn_param_stride(const BnnMlpDesc* d) {
  Layout L;
  if (make_layout(d, &L)) return -1;
  return L.off[L.n_lin];
}
extern "C" int64_t bnn_layer_offset(const BnnMlpDesc* d, int32_t l) {
  Layout L;
  if (make_layout(d, &L)) return -1;
  if (l < 0 || l > L.n_lin) { set_error("layer %d out of range", l); return -1; }
  return L.off[l];
}
extern "C" int32_t bnn_layer_ld(const BnnMlpDesc* d, int32_t l) {
  Layout L;
  if (make_layout(d, &L)) return -1;
  if (l < 0 || l >= L.n_lin) { set_error("layer %d out of range", l); return -1; }
  return L.ld[l];
}
extern "C" int64_t bnn_mask_len(const BnnMlpDesc* d, const int32_t* layer_masked) {
  Layout L;
  if (make_layout(d, &L)) return -1;
  int mo[BNN_MAX_LINEAR];
  return make_mask_offsets(L, layer_masked, mo);
}

extern "C" int64_t bnn_forward_workspace_bytes(const BnnForwardArgs* a) {
  if (!a) { set_error("null args"); return -1; }
  return use_tc(a) ? tc_workspace_bytes(a) : forward_simt_workspace_bytes(a);
}

extern "C" int32_t bnn_forward(const BnnForwardArgs* a, void* stream) {
  if (check_forward_args(a, false)) return -1;
  return use_tc(a) ? tc_forward(a, (cudaStream_t)stream) : forward_simt(a, (cudaStream_t)stream);
}

// ---- host-buffer path ---------------------------------------------------------
namespace {
struct HostWs {
  void* buf[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  cudaStream_t st = nullptr, st_copy = nullptr;
  cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  int ensure(int i, size_t bytes) {
    if (bytes <= cap[i]) return 0;
    if (buf[i]) cudaFree(buf[i]);
    buf[i] = nullptr;
    cap[i] = 0;
    BNN_CUDA(cudaMalloc(&buf[i], bytes));
    cap[i] = bytes;
    return 0;
  }
  void release() {
    for (int i = 0; i < 10; ++i) {
      if (buf[i]) cudaFree(buf[i]);
      buf[i] = nullptr;
      cap[i] = 0;
    }
    for (int i = 0; i < 2; ++i) {
      if (ev_up[i]) cudaEventDestroy(ev_up[i]);
      if (ev_free[i]) cudaEventDestroy(ev_free[i]);
      ev_up[i] = ev_free[i] = nullptr;
    }
    if (st) cudaStreamDestroy(st);
    if (st_copy) cudaStreamDestroy(st_copy);
    st = st_copy = nullptr;
  }
};
// one workspace (buffers, streams, events) per device: the buffers of device 0 must never be handed to a launch on device 1
constexpr int kMaxDevices = 64;
HostWs g_ws_dev[kMaxDevices];
std::mutex g_ws_mu;
}  // namespace

extern "C" void bnn_release_workspace(void) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < kMaxDevices; ++d) {
    bool used = g_ws_dev[d].st || g_ws_dev[d].st_copy;
    for (int i = 0; i < 10; ++i) used = used || g_ws_dev[d].buf[i];
    if (!used) continue;
    cudaSetDevice(d);
    g_ws_dev[d].release();
  }
  cudaSetDevice(cur);
}

extern "C" int32_t bnn_forward_host(const BnnForwardArgs* ah) {
  if (check_forward_args(ah, true)) return -1;
  std::lock_guard<std::mutex> lk(g_ws_mu);
  Layout L;
  if (make_layout(&ah->desc, &L)) return -1;
  int cur_dev = 0;
  BNN_CUDA(cudaGetDevice(&cur_dev));
  if (cur_dev < 0 || cur_dev >= kMaxDevices) BNN_FAIL("bnn_forward_host: device index %d out of range", cur_dev);
  HostWs& g_ws = g_ws_dev[cur_dev];
  if (!g_ws.st) BNN_CUDA(cudaStreamCreateWithFlags(&g_ws.st, cudaStreamNonBlocking));
  cudaStream_t st = g_ws.st;
  BnnForwardArgs a = *ah;
  const int nout = L.out[L.n_lin - 1];
  const long long M = a.N * nout;
  const size_t x_bytes = (size_t)a.N * L.in[0] * 4;
  const long long nets = a.w_stride == 0 ? 1 : a.S;
  const long long wst = a.w_stride == 0 ? L.off[L.n_lin] : a.w_stride;
  const size_t w_bytes = (size_t)nets * wst * 4;
  enum { BX = 0, BW, BMASK, BPRED, BMEAN, BVAR, BPART, BWS, BW2, BCHUNK };
  if (g_ws.ensure(BX, x_bytes)) return -2;
  BNN_CUDA(cudaMemcpyAsync(g_ws.buf[BX], ah->X, x_bytes, cudaMemcpyHostToDevice, st));
  a.X = (const float*)g_ws.buf[BX];
  if (a.mask.mode == BNN_MASK_EXTERNAL) {
    const size_t mb = (size_t)a.S * a.mask.mask_stride * 4;
    if (g_ws.ensure(BMASK, mb)) return -2;
    BNN_CUDA(cudaMemcpyAsync(g_ws.buf[BMASK], ah->mask.mask, mb, cudaMemcpyHostToDevice, st));
    a.mask.mask = (const float*)g_ws.buf[BMASK];
  }
  // ---- large per-sample banks, moments only: upload the bank in chunks on a copy stream behind the evaluation of the
  //      previous chunk; every chunk is reduced to float64 (mean, M2) per point and the chunks are Chan-merged
  //      (bnn_moments_merge), which is the arithmetic of the multi-GPU shard merge
  const long long kChunk = 256;
  if (a.w_stride != 0 && a.S >= 3 * kChunk && !ah->pred && (ah->mean || ah->var || ah->part_mean || ah->part_m2)) {
    long long chunk = kChunk;
    while ((a.S + chunk - 1) / chunk > 64) chunk *= 2;          // bnn_moments_merge takes at most 64 shards
    const int R = (int)((a.S + chunk - 1) / chunk);
    const size_t cb = (size_t)chunk * a.w_stride * 4;
    if (!g_ws.st_copy) BNN_CUDA(cudaStreamCreateWithFlags(&g_ws.st_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      if (!g_ws.ev_up[i]) BNN_CUDA(cudaEventCreateWithFlags(&g_ws.ev_up[i], cudaEventDisableTiming));
      if (!g_ws.ev_free[i]) BNN_CUDA(cudaEventCreateWithFlags(&g_ws.ev_free[i], cudaEventDisableTiming));
    }
    if (g_ws.ensure(BW, cb) || g_ws.ensure(BW2, cb) || g_ws.ensure(BCHUNK, (size_t)R * M * 16)) return -2;
    if (g_ws.ensure(BMEAN, (size_t)M * 4) || g_ws.ensure(BVAR, (size_t)M * 4)) return -2;
    double* parts = (double*)g_ws.buf[BCHUNK];                  // [R][M] means, then [R][M] M2s
    int64_t counts[64];
    BnnForwardArgs c = a;
    c.mean = c.var = nullptr;
    c.S = chunk;
    const long long wsb = use_tc(&c) ? tc_workspace_bytes(&c) : forward_simt_workspace_bytes(&c);
    if (wsb < 0) return -1;
    if (g_ws.ensure(BWS, (size_t)(wsb > 16 ? wsb : 16))) return -2;
    c.workspace = g_ws.buf[BWS];
    c.workspace_bytes = wsb;
    for (int i = 0; i < R; ++i) {
      const long long lo = (long long)i * chunk, n = (a.S - lo < chunk) ? a.S - lo : chunk;
      void* wb = g_ws.buf[(i & 1) ? BW2 : BW];
      if (i >= 2) BNN_CUDA(cudaStreamWaitEvent(g_ws.st_copy, g_ws.ev_free[i & 1], 0));   // chunk i-2 has been evaluated
      BNN_CUDA(cudaMemcpyAsync(wb, ah->W + lo * a.w_stride, (size_t)n * a.w_stride * 4, cudaMemcpyHostToDevice, g_ws.st_copy));
      BNN_CUDA(cudaEventRecord(g_ws.ev_up[i & 1], g_ws.st_copy));
      BNN_CUDA(cudaStreamWaitEvent(st, g_ws.ev_up[i & 1], 0));
      c.W = (const float*)wb;
      c.S = n;
      c.part_mean = parts + (long long)i * M;
      c.part_m2 = parts + (long long)(R + i) * M;
      if (a.mask.mode == BNN_MASK_EXTERNAL) c.mask.mask = a.mask.mask + lo * a.mask.mask_stride;
      c.mask.sample_offset = a.mask.sample_offset + lo;          // in-kernel Philox streams are indexed by the global sample
      counts[i] = n;
      const int rc = use_tc(&c) ? tc_forward(&c, st) : forward_simt(&c, st);
      if (rc) return rc;
      BNN_CUDA(cudaEventRecord(g_ws.ev_free[i & 1], st));
    }
    if (ah->mean || ah->var) {
      if (bnn_moments_merge(parts, parts + (long long)R * M, counts, R, M, ah->mean ? (float*)g_ws.buf[BMEAN] : nullptr,
                            ah->var ? (float*)g_ws.buf[BVAR] : nullptr, st))
        return -2;
    }
    if (ah->part_mean || ah->part_m2) {   // the shard's float64 moments (multi-GPU merge input)
      if (g_ws.ensure(BPART, (size_t)M * 16)) return -2;
      if (bnn_moments_merge_partials(parts, parts + (long long)R * M, counts, R, M, (double*)g_ws.buf[BPART],
                                     (double*)g_ws.buf[BPART] + M, st))
        return -2;
      if (ah->part_mean) BNN_CUDA(cudaMemcpyAsync(ah->part_mean, g_ws.buf[BPART], (size_t)M * 8, cudaMemcpyDeviceToHost, st));
      if (ah->part_m2) BNN_CUDA(cudaMemcpyAsync(ah->part_m2, (double*)g_ws.buf[BPART] + M, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
    }
    if (ah->mean) BNN_CUDA(cudaMemcpyAsync(ah->mean, g_ws.buf[BMEAN], (size_t)M * 4, cudaMemcpyDeviceToHost, st));
    if (ah->var) BNN_CUDA(cudaMemcpyAsync(ah->var, g_ws.buf[BVAR], (size_t)M * 4, cudaMemcpyDeviceToHost, st));
    BNN_CUDA(cudaStreamSynchronize(st));
    return 0;
  }
  if (g_ws.ensure(BW, w_bytes)) return -2;
  BNN_CUDA(cudaMemcpyAsync(g_ws.buf[BW], ah->W, w_bytes, cudaMemcpyHostToDevice, st));
  a.W = (const float*)g_ws.buf[BW];
  if (ah->pred) {
    if (g_ws.ensure(BPRED, (size_t)a.S * M * 4)) return -2;
    a.pred = (float*)g_ws.buf[BPRED];
  }
  if (ah->mean) { if (g_ws.ensure(BMEAN, (size_t)M * 4)) return -2; a.mean = (float*)g_ws.buf[BMEAN]; }
  if (ah->var) { if (g_ws.ensure(BVAR, (size_t)M * 4)) return -2; a.var = (float*)g_ws.buf[BVAR]; }
  if (ah->part_mean || ah->part_m2) {
    if (g_ws.ensure(BPART, (size_t)M * 16)) return -2;
    a.part_mean = ah->part_mean ? (double*)g_ws.buf[BPART] : nullptr;
    a.part_m2 = ah->part_m2 ? (double*)g_ws.buf[BPART] + M : nullptr;
  }
  const long long wsb = use_tc(&a) ? tc_workspace_bytes(&a) : forward_simt_workspace_bytes(&a);
  if (wsb < 0) return -1;
  if (g_ws.ensure(BWS, (size_t)(wsb > 16 ? wsb : 16))) return -2;
  a.workspace = g_ws.buf[BWS];
  a.workspace_bytes = wsb;
  const int rc = use_tc(&a) ? tc_forward(&a, st) : forward_simt(&a, st);
  if (rc) return rc;
  if (ah->pred) BNN_CUDA(cudaMemcpyAsync(ah->pred, a.pred, (size_t)a.S * M * 4, cudaMemcpyDeviceToHost, st));
  if (ah->mean) BNN_CUDA(cudaMemcpyAsync(ah->mean, a.mean, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
  if (ah->var) BNN_CUDA(cudaMemcpyAsync(ah->var, a.var, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
  if (ah->part_mean) BNN_CUDA(cudaMemcpyAsync(ah->part_mean, a.part_mean, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
  if (ah->part_m2) BNN_CUDA(cudaMemcpyAsync(ah->part_m2, a.part_m2, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
  BNN_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---- measurement helper: FP32 FFMA peak of this device (roofline denominator for the SIMT forward)
namespace bnn {
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  const float b = 1.0000001f, c = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678f) out[0] = s;
}
}  // namespace bnn

extern "C" double bnn_measure_fp32_tflops(void) {
  float* d = nullptr;
  if (cudaMalloc(&d, 4) != cudaSuccess) return -1.0;
  const int iters = 4096, blocks = sm_count() * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    ffma_peak_kernel<<<blocks, 256>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * 8 * (double)iters * 256.0 * blocks;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}
