// Probe: issue throughput of FFMA vs FFMA2 (packed fp32x2) on sm_100a.  nvcc -arch=sm_100a -o f32x2_probe f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int MODE> __global__ void k(float* out, float a, float b, int iters) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  } else {
    u64 p[8], pa = pk(a, a), pb = pk(b, b);
    for (int i = 0; i < 8; ++i) p[i] = pk(x[2 * i], x[2 * i + 1]);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
    for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(p[i]));
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int thr : {128, 256, 512, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148, thr>>>(d, 1.0001f, 0.5f, iters); else k<1><<<148, thr>>>(d, 1.0001f, 0.5f, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * thr * 16.0 * iters;
      printf("mode %s threads/SM %4d: %.3f ms  %.1f fp32 FMA/clk/SM (at 1.9 GHz)  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", thr, ms, fma / 148 / (ms * 1e-3 * 1.9e9), 2 * fma / (ms * 1e-3) / 1e12);
    }
  return 0;
}
