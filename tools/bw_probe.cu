// HBM reference points for write-heavy kernels (the GEMM epilogues write 3-8x what they read):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/bw_probe.cu -o tools/bw_probe
// Prints GB/s for: cudaMemset, st.global.v4 streams, ld.global.v4 streams, copy, a 1:4 read:write mix, TMA bulk stores
// with 64-byte / 128-byte / 4-KB contiguous pieces (the piece width is what a tiled TMA store writes per row).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void wr_kernel(int4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_int4(1, 2, 3, 4);
}
__global__ void rd_kernel(const int4* p, size_t n, int* sink) {
  int acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { int4 v = __ldg(p + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678) *sink = acc;
}
__global__ void cp_kernel(const int4* a, int4* b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = __ldg(a + i);
}
// read n/4 vectors, write n vectors (fc1-like 1:4)
__global__ void mix_kernel(const int4* a, int4* b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x) {
    int4 v = __ldg(a + i);
    b[4 * i] = v; b[4 * i + 1] = v; b[4 * i + 2] = v; b[4 * i + 3] = v;
  }
}
// Strided-piece writes: each warp writes 32 rows x `piece` bytes, rows `pitch` bytes apart (what a 32-row TMA box does),
// neighbouring pieces of the same rows are written by OTHER warps.
template <int PIECE>
__global__ void piece_kernel(uint8_t* base, size_t rows, int pitch) {
  constexpr int LPR = PIECE / 16;             // lanes per row
  const int pieces_per_row = pitch / PIECE;
  const size_t row_blocks = rows / 32;
  const size_t total = row_blocks * pieces_per_row;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (size_t t = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5; t < total; t += warps) {
    const size_t rb = t / pieces_per_row; const int pc = (int)(t % pieces_per_row);
#pragma unroll
    for (int k = 0; k < 32 * LPR / 32; ++k) {
      const int idx = k * 32 + lane;
      const int r = idx / LPR, c = idx % LPR;
      *reinterpret_cast<int4*>(base + (rb * 32 + r) * (size_t)pitch + pc * PIECE + c * 16) = make_int4(1, 2, 3, 4);
    }
  }
}

template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < 5; ++i) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best * 1e-3f;
}

int main() {
  const size_t bytes = 2ull << 30, n = bytes / 16;
  int4 *a, *b; int* sink;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 4);
  cudaMemset(a, 1, bytes); cudaMemset(b, 1, bytes);
  const int grid = 148 * 8, blk = 512;
  printf("cudaMemset        %7.0f GB/s\n", bytes / timeit([&] { cudaMemsetAsync(a, 0, bytes); }) / 1e9);
  printf("st.v4 stream      %7.0f GB/s\n", bytes / timeit([&] { wr_kernel<<<grid, blk>>>(a, n); }) / 1e9);
  printf("ld.v4 stream      %7.0f GB/s\n", bytes / timeit([&] { rd_kernel<<<grid, blk>>>(a, n, sink); }) / 1e9);
  printf("copy              %7.0f GB/s (read+write)\n", 2.0 * bytes / timeit([&] { cp_kernel<<<grid, blk>>>(a, b, n); }) / 1e9);
  printf("mix 1r:4w         %7.0f GB/s (read+write)\n", 1.25 * bytes / timeit([&] { mix_kernel<<<grid, blk>>>(a, b, n); }) / 1e9);
  const int pitch = 768;                       // e.g. a (rows, 384) bf16 matrix
  const size_t rows = bytes / pitch / 32 * 32;
  printf("pieces  64B/row   %7.0f GB/s\n", rows * (double)pitch / timeit([&] { piece_kernel<64><<<grid, blk>>>((uint8_t*)a, rows, pitch); }) / 1e9);
  printf("pieces 128B/row   %7.0f GB/s\n", rows * (double)pitch / timeit([&] { piece_kernel<128><<<grid, blk>>>((uint8_t*)a, rows, pitch); }) / 1e9);
  printf("pieces 256B/row   %7.0f GB/s\n", rows * (double)pitch / timeit([&] { piece_kernel<256><<<grid, blk>>>((uint8_t*)a, rows, pitch); }) / 1e9);
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
