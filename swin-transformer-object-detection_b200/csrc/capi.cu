// C-ABI front door: version / error reporting / device check and dtype dispatch.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <set>
#include <utility>

#include <stdlib.h>
#include "common.cuh"

namespace swin {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<int> g_sm_count[64];       // 0 = not queried yet
static std::atomic<int> g_sm_reserve{0};
int persistent_sms() {
  int dev = 0, n = kNumSMs;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
    n = g_sm_count[dev].load(std::memory_order_relaxed);
    if (n == 0) {
      if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
      g_sm_count[dev].store(n, std::memory_order_relaxed);
    }
  } else {
    (void)cudaGetLastError();                 // no device: keep the sticky-error state clean for the host-only entry points
  }
  n -= g_sm_reserve.load(std::memory_order_relaxed);
  n &= ~1;                                    // CTA pairs occupy whole TPCs
  return n < 2 ? 2 : n;
}

static std::mutex g_attr_mu;
static std::set<std::pair<const void*, int>> g_attr_done;
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("SWIN_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
int ensure_dyn_smem(const void* func, int bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_attr_mu);
  if (g_attr_done.count({func, dev})) return 0;
  if (bytes <= 0) {                             // "as much as this kernel can have": opt-in limit minus its static shared memory
    cudaFuncAttributes fa;
    int optin = 0;
    cudaError_t e0 = cudaFuncGetAttributes(&fa, func);
    if (e0 == cudaSuccess) e0 = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e0 != cudaSuccess) { set_error("cudaFuncGetAttributes: %s", cudaGetErrorString(e0)); return (int)e0; }
    bytes = optin - (int)fa.sharedSizeBytes;
  }
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  g_attr_done.insert({func, dev});
  return 0;
}
}  // namespace swin

using namespace swin;

extern "C" int swin_sm_reserve(int n) {
  if (n < 0) return g_sm_reserve.load(std::memory_order_relaxed);
  return g_sm_reserve.exchange(n > 64 ? 64 : n, std::memory_order_relaxed);
}

extern "C" int swin_version(void) { return SWIN_B200_VERSION; }
extern "C" const char* swin_last_error(void) { return g_err; }

extern "C" int swin_device_check(int device) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) { set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  if (major != 10) { set_error("device %d has compute capability %d.x; these kernels are sm_100a only", device, major); return -ENOTSUP; }
  return 0;
}

extern "C" int swin_gemm(const swin_gemm_args* a, void* stream) {
  if (!a) { set_error("gemm: null args"); return -EINVAL; }
  if (a->dtype == SWIN_F32) return gemm_simt(a, (cudaStream_t)stream);
  if (a->dtype == SWIN_BF16) return gemm_tc(a, (cudaStream_t)stream);
  set_error("gemm: bad dtype %d", a->dtype);
  return -EINVAL;
}
extern "C" int swin_gemm_pair_mode(int mode) { return set_pair_mode(mode); }
extern "C" int swin_gemm_plan(const swin_gemm_args* a, int* out6) { return gemm_tc_plan(a, out6); }
extern "C" int swin_window_attn_fwd(const swin_attn_args* a, void* stream) {
  if (!a) { set_error("attn: null args"); return -EINVAL; }
  if (a->dtype == SWIN_F32) return attn_simt_fwd(a, (cudaStream_t)stream);
  if (a->dtype == SWIN_BF16) return a->ws == 7 ? attn_tc_fwd(a, (cudaStream_t)stream) : (a->ws == 12 ? attn_mma_fwd(a, (cudaStream_t)stream) : attn_simt_fwd(a, (cudaStream_t)stream));
  set_error("attn: bad dtype %d", a->dtype);
  return -EINVAL;
}
extern "C" int swin_window_attn_bwd(const swin_attn_args* a, void* stream) {
  if (!a) { set_error("attn: null args"); return -EINVAL; }
  if (a->dtype == SWIN_F32) return attn_simt_bwd(a, (cudaStream_t)stream);
  if (a->dtype == SWIN_BF16) return a->ws == 7 ? attn_tc_bwd(a, (cudaStream_t)stream) : (a->ws == 12 ? attn_mma_bwd(a, (cudaStream_t)stream) : attn_simt_bwd(a, (cudaStream_t)stream));
  set_error("attn: bad dtype %d", a->dtype);
  return -EINVAL;
}
extern "C" int swin_window_attn_qkv_fwd(const swin_attn_qkv_args* a, void* stream) {
  if (!a) { set_error("attn_qkv: null args"); return -EINVAL; }
  return attn_qkv_fwd(a, (cudaStream_t)stream);
}
extern "C" int swin_window_attn_qkv_supported(int C, int nH, int ws) { return attn_qkv_supported(C, nH, ws); }
extern "C" long long swin_window_attn_qkv_workspace(int C, int nH, int ws) { return attn_qkv_workspace_bytes(C, nH, ws); }
