// Shared host/device helpers for the swin_b200 kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/swin_b200.h"

namespace swin {

constexpr int kNumSMs = 148;    // B200; grid-size heuristic of the (over-subscribed) HBM-bound kernels only

void set_error(const char* fmt, ...);
// SMs of the CURRENT device (queried once per device; 148 when no device is visible, so the host-only tile planner works
// without a GPU) minus the SMs reserved with swin_sm_reserve(): the grid size of the persistent kernels (GEMM, attention).
int persistent_sms();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device); thread-safe (forward runs on the main thread,
// backward on autograd's worker).  Returns 0 or a cudaError_t (message in swin_last_error()).
int ensure_dyn_smem(const void* func, int bytes);

#define SWIN_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      ::swin::set_error(__VA_ARGS__);  \
      return -EINVAL;                  \
    }                                  \
  } while (0)

#define SWIN_LAUNCH_CHECK()                                                    \
  do {                                                                         \
    cudaError_t e__ = cudaGetLastError();                                      \
    if (e__ != cudaSuccess) {                                                  \
      ::swin::set_error("%s:%d launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return (int)e__;                                                         \
    }                                                                          \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Division by a launch-time constant as multiply-high + shift (valid for 0 <= n < 2^31): the gather / scatter index
// math runs once per row in HBM-bound kernels whose issue slots, not their bytes, were the limit (ncu: 75 % issue-slot
// utilisation at 64 % of HBM peak with hardware-emulated integer division).
struct FastDiv {
  uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = (uint32_t)(d < 1 ? 1 : d); f.mul = 0; f.shr = 0;
  if (f.d > 1) {
    int l = 0;
    while ((1u << l) < f.d) ++l;                       // ceil(log2 d)
    const int p = 31 + l;
    f.mul = (uint32_t)((((uint64_t)1 << p) + f.d - 1) / f.d);
    f.shr = (uint32_t)(p - 32);
  }
  return f;
}
__host__ __device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
#ifdef __CUDA_ARCH__
  return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
#else
  return n / (int)f.d;
#endif
}

// Geometry of one stage's window grid (REF:214-231 / :371-372).
struct WinGeom {
  int B, H, W, C, ws, shift;
  int Hp, Wp, nwh, nww, nW, N;
  FastDiv dN, dnww, dws, dW, dslots, dtok;     // dividers: N, nww, ws, W, nW*N (slots per image), H*W (tokens per image)
};
inline WinGeom make_geom(int B, int H, int W, int C, int ws, int shift) {
  WinGeom g;
  g.B = B; g.H = H; g.W = W; g.C = C; g.ws = ws; g.shift = shift;
  g.nwh = (H + ws - 1) / ws; g.nww = (W + ws - 1) / ws;
  g.Hp = g.nwh * ws; g.Wp = g.nww * ws;
  g.nW = g.nwh * g.nww; g.N = ws * ws;
  g.dN = make_fastdiv(g.N); g.dnww = make_fastdiv(g.nww); g.dws = make_fastdiv(ws); g.dW = make_fastdiv(W);
  g.dslots = make_fastdiv(g.nW * g.N); g.dtok = make_fastdiv(H * W);
  return g;
}

// slot (window-token index inside ONE image, [0, nW*N)) -> source token h*W+w, or -1 for zero padding.
__host__ __device__ __forceinline__ int slot_to_token(const WinGeom& g, int slot) {
  int w = fdiv(slot, g.dN), t = slot - w * g.N;
  int wh = fdiv(w, g.dnww), ww = w - wh * g.nww;
  int i = fdiv(t, g.dws), j = t - i * g.ws;
  int hs = wh * g.ws + i + g.shift; if (hs >= g.Hp) hs -= g.Hp;
  int wsrc = ww * g.ws + j + g.shift; if (wsrc >= g.Wp) wsrc -= g.Wp;
  return (hs < g.H && wsrc < g.W) ? hs * g.W + wsrc : -1;
}
// token (h*W+w) -> slot inside the image (inverse of the above on valid tokens).
__host__ __device__ __forceinline__ int token_to_slot(const WinGeom& g, int tok) {
  int h = fdiv(tok, g.dW), w = tok - h * g.W;
  int hh = h - g.shift; if (hh < 0) hh += g.Hp;
  int wq = w - g.shift; if (wq < 0) wq += g.Wp;
  int wh = fdiv(hh, g.dws), i = hh - wh * g.ws;
  int ww = fdiv(wq, g.dws), j = wq - ww * g.ws;
  return (wh * g.nww + ww) * g.N + i * g.ws + j;
}
// global row (image-major) -> (image, row inside the image) for window-slot rows / token rows
__device__ __forceinline__ int split_slot_row(const WinGeom& g, int row, int* inner) {
  const int b = fdiv(row, g.dslots); *inner = row - b * (g.nW * g.N); return b;
}
__device__ __forceinline__ int split_tok_row(const WinGeom& g, int row, int* inner) {
  const int b = fdiv(row, g.dtok); *inner = row - b * (g.H * g.W); return b;
}

__device__ __forceinline__ float gelu_erf(float u) { return 0.5f * u * (1.0f + erff(u * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_erf(float u) {
  // d/du [u * Phi(u)] = Phi(u) + u * phi(u)
  float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752440f));
  float pdf = 0.39894228040143267794f * __expf(-0.5f * u * u);
  return cdf + u * pdf;
}

// Fast GELU / GELU' for the bf16 tensor-core epilogues (the fp32 parity mode uses erff): erfc by Abramowitz-Stegun
// 7.1.25 (|err| <= 2.5e-5, two orders below bf16 rounding), constants pre-folded so one element costs
// 1 MUFU.RCP + 1 MUFU.EX2 + ~12 FMA-pipe ops; exp(-u^2/2) is shared by the cdf and the pdf.
__device__ __forceinline__ void gelu_fast(float u, float* g, float* dg) {
  float t, E;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.33267111f, fabsf(u), 1.0f)));          // 1 / (1 + p |u| / sqrt2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E) : "f"(u * (-0.72134752044448170368f * u)));          // exp(-u^2/2)
  const float q = fmaf(fmaf(0.3739278f, t, -0.0479399f), t, 0.1740121f);                           // 0.5 * (a1 + a2 t + a3 t^2)
  const float half_erfc = q * t * E;                                                               // 0.5 * erfc(|u|/sqrt2)
  const float cdf = u >= 0.f ? 1.0f - half_erfc : half_erfc;
  *g = u * cdf;
  *dg = fmaf(u * 0.39894228040143267794f, E, cdf);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Programmatic dependent launch (PDL): a kernel launched with the attribute may begin while its predecessor in the stream is still
// running; everything it does before pdl_wait() (barrier init, TMEM allocation, shared-memory zeroing: no global access) overlaps
// the predecessor's tail.  pdl_wait() returns once every prerequisite grid has completed and its writes are visible; pdl_trigger()
// (issued right after it, so completion stays transitive along the stream) lets the NEXT kernel's CTAs be scheduled as SMs free up.
// Under stream capture the attribute becomes a programmatic edge of the CUDA graph.  SWIN_PDL=0 launches everything plainly.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
// fills `attr` with the programmatic-serialization attribute; returns 1 if PDL is on (number of attributes written)
inline int pdl_attr(cudaLaunchAttribute* attr) {
  if (!pdl_enabled()) return 0;
  attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr->val.programmaticStreamSerializationAllowed = 1;
  return 1;
}

// internal entry points implemented per translation unit
int gemm_simt(const swin_gemm_args* a, cudaStream_t st);
int gemm_tc(const swin_gemm_args* a, cudaStream_t st);
int set_pair_mode(int mode);
int gemm_tc_plan(const swin_gemm_args* a, int* out6);
int attn_simt_fwd(const swin_attn_args* a, cudaStream_t st);
int attn_simt_bwd(const swin_attn_args* a, cudaStream_t st);
int attn_mma_fwd(const swin_attn_args* a, cudaStream_t st);
int attn_mma_bwd(const swin_attn_args* a, cudaStream_t st);
int attn_tc_fwd(const swin_attn_args* a, cudaStream_t st);
int attn_tc_bwd(const swin_attn_args* a, cudaStream_t st);
int attn_qkv_fwd(const swin_attn_qkv_args* a, cudaStream_t st);
int attn_qkv_supported(int C, int nH, int ws);
long long attn_qkv_workspace_bytes(int C, int nH, int ws);

}  // namespace swin
