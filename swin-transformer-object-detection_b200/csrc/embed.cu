// The two ends of the block stack:
//   * PatchEmbed (REF:429-445): zero-pad + 4x4/4 conv as a patch-unfold gather feeding the GEMM kernels;
//   * per-stage output norm + NHWC->NCHW (REF:618-623): LayerNorm whose store is transposed through smem,
//     and the matching backward that reads the NCHW gradient through the same transpose.
#include "common.cuh"

namespace swin {

// ------------------------------------------------------------------------------------------ patch unfold
// img (B, Cin, Hi, Wi) fp32  <->  cols (B*Hh*Ww, Cin*p*p), column order [c][i][j] == weight.view(C, -1)
template <typename T, bool SCATTER>
__global__ void __launch_bounds__(256) patch_unfold_kernel(float* __restrict__ img, T* __restrict__ cols, int B, int Cin, int Hi,
                                                           int Wi, int p, int Hh, int Ww) {
  const long long total = (long long)B * Cin * Hh * p * Ww;
  const int K = Cin * p * p;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(e % Ww);
    long long r = e / Ww;
    const int y = (int)(r % (Hh * p)); r /= (Hh * p);
    const int c = (int)(r % Cin);
    const int b = (int)(r / Cin);
    const int ph = y / p, i = y - ph * p;
    T* dst = cols + ((long long)(b * Hh + ph) * Ww + pw) * K + c * p * p + i * p;
    float* src = img + (((long long)b * Cin + c) * Hi + y) * Wi + (long long)pw * p;
    for (int j = 0; j < p; ++j) {
      const bool in = (y < Hi) && (pw * p + j < Wi);
      if (SCATTER) { if (in) src[j] = (float)dst[j]; }
      else dst[j] = (T)(in ? src[j] : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------ LN + NCHW store
constexpr int kTokTile = 32;
constexpr int kMaxVpl = 32;      // C <= 1024

__global__ void __launch_bounds__(256) ln_nchw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ out,
                                                          float* __restrict__ mean, float* __restrict__ rstd, int L, int C, float eps) {
  extern __shared__ float tile[];            // [C][33]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y, t0 = blockIdx.x * kTokTile;
  const int vpl = (C + 31) / 32;
  const float inv_n = 1.0f / (float)C;
  for (int tt = warp; tt < kTokTile; tt += 8) {
    const int t = t0 + tt;
    if (t >= L) break;
    const float* row = x + ((long long)b * L + t) * C;
    float r[kMaxVpl];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxVpl; ++k) {
      r[k] = 0.f;
      if (k < vpl && lane + 32 * k < C) { r[k] = __ldg(row + lane + 32 * k); s += r[k]; }
    }
    const float mu = warp_sum(s) * inv_n;
    float q2 = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxVpl; ++k)
      if (k < vpl && lane + 32 * k < C) { float d = r[k] - mu; q2 += d * d; }
    const float rs = rsqrtf(warp_sum(q2) * inv_n + eps);
    if (lane == 0) { mean[(long long)b * L + t] = mu; rstd[(long long)b * L + t] = rs; }
#pragma unroll
    for (int k = 0; k < kMaxVpl; ++k) {
      const int c = lane + 32 * k;
      if (k < vpl && c < C) tile[c * 33 + tt] = (r[k] - mu) * rs * __ldg(gamma + c) + __ldg(beta + c);
    }
  }
  __syncthreads();
  const int t = t0 + lane;
  if (t < L)
    for (int c = warp; c < C; c += 8) out[((long long)b * C + c) * L + t] = tile[c * 33 + lane];
}

__global__ void __launch_bounds__(256) ln_nchw_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                          const float* __restrict__ gamma, const float* __restrict__ mean,
                                                          const float* __restrict__ rstd, float* __restrict__ dx,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, int L, int C,
                                                          int tiles_per_block) {
  extern __shared__ float tile[];            // [C][33] + [2][C] partials
  float* sred = tile + (size_t)C * 33;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int vpl = (C + 31) / 32;
  const float inv_n = 1.0f / (float)C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;
  float ag[kMaxVpl], ab[kMaxVpl];
#pragma unroll
  for (int k = 0; k < kMaxVpl; ++k) { ag[k] = 0.f; ab[k] = 0.f; }
  for (int tb = 0; tb < tiles_per_block; ++tb) {
    const int t0 = (blockIdx.x * tiles_per_block + tb) * kTokTile;
    if (t0 >= L) break;
    __syncthreads();
    {
      const int t = t0 + lane;
      for (int c = warp; c < C; c += 8) tile[c * 33 + lane] = (t < L) ? __ldg(dout + ((long long)b * C + c) * L + t) : 0.f;
    }
    __syncthreads();
    for (int tt = warp; tt < kTokTile; tt += 8) {
      const int t = t0 + tt;
      if (t >= L) break;
      const long long rowi = (long long)b * L + t;
      const float mu = mean[rowi], rs = rstd[rowi];
      float xh[kMaxVpl], gd[kMaxVpl];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxVpl; ++k) {
        const int c = lane + 32 * k;
        xh[k] = 0.f; gd[k] = 0.f;
        if (k < vpl && c < C) {
          const float d = tile[c * 33 + tt];
          xh[k] = (__ldg(x + rowi * C + c) - mu) * rs;
          gd[k] = d * __ldg(gamma + c);
          s1 += gd[k]; s2 += gd[k] * xh[k];
          ag[k] += d * xh[k]; ab[k] += d;
        }
      }
      const float m1 = warp_sum(s1) * inv_n, m2 = warp_sum(s2) * inv_n;
#pragma unroll
      for (int k = 0; k < kMaxVpl; ++k) {
        const int c = lane + 32 * k;
        if (k < vpl && c < C) dx[rowi * C + c] = rs * (gd[k] - m1 - xh[k] * m2);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kMaxVpl; ++k) {
    const int c = lane + 32 * k;
    if (k < vpl && c < C) { atomicAdd(&sred[c], ag[k]); atomicAdd(&sred[C + c], ab[k]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) { atomicAdd(dgamma + i, sred[i]); atomicAdd(dbeta + i, sred[C + i]); }
}

}  // namespace swin

using namespace swin;

static int patch_geom_check(int B, int Cin, int Hi, int Wi, int p) {
  SWIN_REQUIRE(B > 0 && Cin > 0 && Hi > 0 && Wi > 0 && p > 0 && p <= 16, "patch: bad shape");
  return 0;
}

extern "C" int swin_patch_gather(const float* img, void* cols, int B, int Cin, int Hi, int Wi, int patch, int dtype, void* stream) {
  int rc = patch_geom_check(B, Cin, Hi, Wi, patch);
  if (rc) return rc;
  const int Hh = ceil_div(Hi, patch), Ww = ceil_div(Wi, patch);
  long long total = (long long)B * Cin * Hh * patch * Ww;
  int grid = (int)(ceil_div64(total, 256) < (long long)kNumSMs * 16 ? ceil_div64(total, 256) : (long long)kNumSMs * 16);
  if (dtype == SWIN_F32) patch_unfold_kernel<float, false><<<grid, 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(img), (float*)cols, B, Cin, Hi, Wi, patch, Hh, Ww);
  else if (dtype == SWIN_BF16) patch_unfold_kernel<__nv_bfloat16, false><<<grid, 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(img), (__nv_bfloat16*)cols, B, Cin, Hi, Wi, patch, Hh, Ww);
  else { set_error("patch_gather: bad dtype"); return -EINVAL; }
  SWIN_LAUNCH_CHECK();
  return 0;
}
extern "C" int swin_patch_scatter(const void* dcols, float* dimg, int B, int Cin, int Hi, int Wi, int patch, int dtype, void* stream) {
  int rc = patch_geom_check(B, Cin, Hi, Wi, patch);
  if (rc) return rc;
  const int Hh = ceil_div(Hi, patch), Ww = ceil_div(Wi, patch);
  long long total = (long long)B * Cin * Hh * patch * Ww;
  int grid = (int)(ceil_div64(total, 256) < (long long)kNumSMs * 16 ? ceil_div64(total, 256) : (long long)kNumSMs * 16);
  if (dtype == SWIN_F32) patch_unfold_kernel<float, true><<<grid, 256, 0, (cudaStream_t)stream>>>(dimg, (float*)const_cast<void*>(dcols), B, Cin, Hi, Wi, patch, Hh, Ww);
  else if (dtype == SWIN_BF16) patch_unfold_kernel<__nv_bfloat16, true><<<grid, 256, 0, (cudaStream_t)stream>>>(dimg, (__nv_bfloat16*)const_cast<void*>(dcols), B, Cin, Hi, Wi, patch, Hh, Ww);
  else { set_error("patch_scatter: bad dtype"); return -EINVAL; }
  SWIN_LAUNCH_CHECK();
  return 0;
}

extern "C" int swin_ln_nchw_fwd(const float* x, const float* gamma, const float* beta, float* out, float* mean, float* rstd, int B,
                                int L, int C, float eps, void* stream) {
  SWIN_REQUIRE(B > 0 && L > 0 && C > 0 && C <= 32 * kMaxVpl, "ln_nchw: bad shape (C <= 1024)");
  SWIN_REQUIRE(x && gamma && beta && out && mean && rstd, "ln_nchw: null pointer");
  size_t smem = (size_t)C * 33 * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(ln_nchw_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; }
  dim3 grid(ceil_div(L, kTokTile), B);
  ln_nchw_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, gamma, beta, out, mean, rstd, L, C, eps);
  SWIN_LAUNCH_CHECK();
  return 0;
}
extern "C" int swin_ln_nchw_bwd(const float* dout, const float* x, const float* gamma, const float* mean, const float* rstd,
                                float* dx, float* dgamma, float* dbeta, int B, int L, int C, void* stream) {
  SWIN_REQUIRE(B > 0 && L > 0 && C > 0 && C <= 32 * kMaxVpl, "ln_nchw: bad shape (C <= 1024)");
  SWIN_REQUIRE(dout && x && gamma && mean && rstd && dx && dgamma && dbeta, "ln_nchw_bwd: null pointer");
  size_t smem = ((size_t)C * 33 + 2 * C) * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(ln_nchw_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; }
  int tiles = ceil_div(L, kTokTile);
  int tpb = ceil_div(tiles * B, kNumSMs * 4);      // a few tiles per block so the dgamma/dbeta atomics amortise
  if (tpb < 1) tpb = 1;
  dim3 grid(ceil_div(tiles, tpb), B);
  ln_nchw_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(dout, x, gamma, mean, rstd, dx, dgamma, dbeta, L, C, tpb);
  SWIN_LAUNCH_CHECK();
  return 0;
}
