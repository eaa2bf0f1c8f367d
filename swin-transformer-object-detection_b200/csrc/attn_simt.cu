// fp32-arithmetic window attention core (FFMA), any window size.  One CTA per (window, head), one thread per query row;
// S/P live in shared memory.  REF:129-150 forward, SURVEY App. E backward.  Two uses: T = float is the <=1e-4 parity mode;
// T = bf16 (bf16 q/k/v/out/dout/dqkv in HBM, fp32 math) serves the window sizes the tcgen05 kernels do not cover
// (window 12 of the 384-pixel Swin-B/L configs), so those models run in bf16 mode with every GEMM on the tensor cores.
#include "common.cuh"

namespace swin {

struct AttnSimtParams {
  int B_, nH, N, nW, C;
  float scale;
  const void* qkv; const float* bias; const float* mask;
  void* out; float* lse;
  const void* dout; void* dqkv; float* dbias;
};
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int HD = 32;

template <typename T>
__global__ void attn_simt_fwd_kernel(AttnSimtParams p) {
  extern __shared__ float sm[];
  const int N = p.N, b = blockIdx.x / p.nH, h = blockIdx.x % p.nH;
  float* sk = sm;                 // [N][33]
  float* sv = sk + N * 33;        // [N][33]
  const T* base = reinterpret_cast<const T*>(p.qkv) + (size_t)b * N * 3 * p.C + h * HD;
  for (int e = threadIdx.x; e < N * HD; e += blockDim.x) {
    int j = e / HD, d = e % HD;
    sk[j * 33 + d] = ldf(base + (size_t)j * 3 * p.C + p.C + d);
    sv[j * 33 + d] = ldf(base + (size_t)j * 3 * p.C + 2 * p.C + d);
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= N) return;
  float q[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) q[d] = ldf(base + (size_t)i * 3 * p.C + d) * p.scale;
  const float* brow = p.bias + ((size_t)h * N + i) * N;
  const float* mrow = p.mask ? p.mask + ((size_t)(b % p.nW) * N + i) * N : nullptr;
  float mx = -INFINITY;
  for (int j = 0; j < N; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s = fmaf(q[d], sk[j * 33 + d], s);
    s += brow[j];
    if (mrow) s += mrow[j];
    mx = fmaxf(mx, s);
  }
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  float sum = 0.f;
  for (int j = 0; j < N; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s = fmaf(q[d], sk[j * 33 + d], s);
    s += brow[j];
    if (mrow) s += mrow[j];
    float e = expf(s - mx);
    sum += e;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = fmaf(e, sv[j * 33 + d], o[d]);
  }
  const float inv = 1.0f / sum;
  T* orow = reinterpret_cast<T*>(p.out) + ((size_t)b * N + i) * p.C + h * HD;
#pragma unroll
  for (int d = 0; d < HD; ++d) stf(orow + d, o[d] * inv);
  p.lse[((size_t)b * p.nH + h) * N + i] = mx + logf(sum);
}

template <typename T>
__global__ void attn_simt_bwd_kernel(AttnSimtParams p) {
  extern __shared__ float sm[];
  const int N = p.N, b = blockIdx.x / p.nH, h = blockIdx.x % p.nH;
  float* sq = sm;                  // [N][33]  (unscaled q)
  float* sk = sq + N * 33;
  float* sv = sk + N * 33;
  float* sdo = sv + N * 33;
  float* sp = sdo + N * 33;        // [N][N+1]  P; dS is recomputed from P, dO, V and delta (keeps window 12 within smem)
  float* sdelta = sp + N * (N + 1);// [N]
  const T* base = reinterpret_cast<const T*>(p.qkv) + (size_t)b * N * 3 * p.C + h * HD;
  const T* dobase = reinterpret_cast<const T*>(p.dout) + (size_t)b * N * p.C + h * HD;
  for (int e = threadIdx.x; e < N * HD; e += blockDim.x) {
    int j = e / HD, d = e % HD;
    sq[j * 33 + d] = ldf(base + (size_t)j * 3 * p.C + d);
    sk[j * 33 + d] = ldf(base + (size_t)j * 3 * p.C + p.C + d);
    sv[j * 33 + d] = ldf(base + (size_t)j * 3 * p.C + 2 * p.C + d);
    sdo[j * 33 + d] = ldf(dobase + (size_t)j * p.C + d);
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i < N) {
    const float* brow = p.bias + ((size_t)h * N + i) * N;
    const float* mrow = p.mask ? p.mask + ((size_t)(b % p.nW) * N + i) * N : nullptr;
    const float l = p.lse[((size_t)b * p.nH + h) * N + i];
    float delta = 0.f;
    for (int j = 0; j < N; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        s = fmaf(sq[i * 33 + d] * p.scale, sk[j * 33 + d], s);
        dp = fmaf(sdo[i * 33 + d], sv[j * 33 + d], dp);
      }
      s += brow[j];
      if (mrow) s += mrow[j];
      float pr = expf(s - l);
      sp[i * (N + 1) + j] = pr;
      delta = fmaf(pr, dp, delta);
    }
    sdelta[i] = delta;
    float dq[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = 0.f;
    for (int j = 0; j < N; ++j) {
      float dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) dp = fmaf(sdo[i * 33 + d], sv[j * 33 + d], dp);
      float ds = sp[i * (N + 1) + j] * (dp - delta);
      atomicAdd(p.dbias + ((size_t)h * N + i) * N + j, ds);
#pragma unroll
      for (int d = 0; d < HD; ++d) dq[d] = fmaf(ds, sk[j * 33 + d], dq[d]);
    }
    T* dqrow = reinterpret_cast<T*>(p.dqkv) + ((size_t)b * N + i) * 3 * p.C + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) stf(dqrow + d, dq[d] * p.scale);
  }
  __syncthreads();
  if (i < N) {
    const int j = i;
    float dk[HD], dv[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    for (int r = 0; r < N; ++r) {
      float dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) dp = fmaf(sdo[r * 33 + d], sv[j * 33 + d], dp);
      const float pr = sp[r * (N + 1) + j];
      const float ds = pr * (dp - sdelta[r]);
#pragma unroll
      for (int d = 0; d < HD; ++d) {
        dk[d] = fmaf(ds, sq[r * 33 + d], dk[d]);
        dv[d] = fmaf(pr, sdo[r * 33 + d], dv[d]);
      }
    }
    T* row = reinterpret_cast<T*>(p.dqkv) + ((size_t)b * N + j) * 3 * p.C + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) { stf(row + p.C + d, dk[d] * p.scale); stf(row + 2 * p.C + d, dv[d]); }
  }
}

static int attn_simt_common(const swin_attn_args* a, AttnSimtParams* out, bool bwd) {
  SWIN_REQUIRE(a->B_ >= 0 && a->nH > 0 && a->ws > 0, "attn: bad shape");
  SWIN_REQUIRE(a->qkv && a->bias && a->lse, "attn: null pointer");
  SWIN_REQUIRE(a->mask == nullptr || (a->nW > 0 && a->B_ % a->nW == 0), "attn: B_ must be a multiple of nW when a mask is given");
  if (bwd) SWIN_REQUIRE(a->dout && a->dqkv && a->dbias, "attn_bwd: null pointer");
  else SWIN_REQUIRE(a->out != nullptr, "attn_fwd: null out");
  AttnSimtParams p;
  p.B_ = a->B_; p.nH = a->nH; p.N = a->ws * a->ws; p.nW = a->nW > 0 ? a->nW : 1; p.C = a->nH * HD; p.scale = a->scale;
  p.qkv = a->qkv; p.bias = a->bias; p.mask = a->mask; p.out = a->out; p.lse = a->lse;
  p.dout = a->dout; p.dqkv = a->dqkv; p.dbias = a->dbias;
  *out = p;
  return 0;
}

int attn_simt_fwd(const swin_attn_args* a, cudaStream_t st) {
  AttnSimtParams p;
  int rc = attn_simt_common(a, &p, false);
  if (rc) return rc;
  if (p.B_ == 0) return 0;
  size_t smem = (size_t)2 * p.N * 33 * sizeof(float);
  int threads = ceil_div(p.N, 32) * 32;
  if (a->dtype == SWIN_BF16) {
    cudaFuncSetAttribute(attn_simt_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attn_simt_fwd_kernel<__nv_bfloat16><<<p.B_ * p.nH, threads, smem, st>>>(p);
  } else {
    cudaFuncSetAttribute(attn_simt_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attn_simt_fwd_kernel<float><<<p.B_ * p.nH, threads, smem, st>>>(p);
  }
  SWIN_LAUNCH_CHECK();
  return 0;
}
int attn_simt_bwd(const swin_attn_args* a, cudaStream_t st) {
  AttnSimtParams p;
  int rc = attn_simt_common(a, &p, true);
  if (rc) return rc;
  if (p.B_ == 0) return 0;
  size_t smem = ((size_t)4 * p.N * 33 + (size_t)p.N * (p.N + 1) + p.N) * sizeof(float);
  SWIN_REQUIRE(smem <= 200 * 1024, "attn_bwd(fp32): window too large for shared memory");
  int threads = ceil_div(p.N, 32) * 32;
  if (a->dtype == SWIN_BF16) {
    cudaFuncSetAttribute(attn_simt_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attn_simt_bwd_kernel<__nv_bfloat16><<<p.B_ * p.nH, threads, smem, st>>>(p);
  } else {
    cudaFuncSetAttribute(attn_simt_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attn_simt_bwd_kernel<float><<<p.B_ * p.nH, threads, smem, st>>>(p);
  }
  SWIN_LAUNCH_CHECK();
  return 0;
}

}  // namespace swin
