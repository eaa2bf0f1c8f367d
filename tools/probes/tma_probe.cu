// Probe: achievable TMA load (+ store) bandwidth for boxes of 49 rows x {64, 128} bytes (the attention kernels' access pattern)
// as a function of CTAs per SM and boxes in flight per CTA.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m), "r"(src), "r"(c0), "r"(c1) : "memory");
}

// one thread per CTA drives a ring of DEPTH stages; a stage = BOXES boxes of 49 rows x WB bytes (+ optionally one store per 2 loads)
template <int WB>
__global__ void probe(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, int depth, int boxes, int nstage_total,
                      int cols_per_row, int store_every) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[16];
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const uint32_t box_bytes = 49 * WB, slot = 64 * WB;      // 64-row slots
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int boxes_per_row = cols_per_row / (WB / 2);
    auto issue = [&](int s) {
      const int st = s % depth;
      const uint32_t bar = smem_u32(&bars[st]);
      mbar_expect(bar, boxes * box_bytes);
      for (int b = 0; b < boxes; ++b) {
        const long long id = ((long long)s * ncta + cta) * boxes + b;       // box id -> (row block, column block): neighbours in a row go to the same CTA
        const int cb = (int)(id % boxes_per_row); const long long rb = id / boxes_per_row;
        tma_load(smem_u32(base + (st * boxes + b) * slot), &tin, bar, cb * (WB / 2), (int)(rb * 49));
      }
    };
    for (int s = 0; s < depth && s < nstage_total; ++s) issue(s);
    for (int s = 0; s < nstage_total; ++s) {
      const int st = s % depth;
      mbar_wait(smem_u32(&bars[st]), (s / depth) & 1);
      if (store_every) {
        for (int b = 0; b < boxes; b += store_every) {
          const long long id = ((long long)s * ncta + cta) * boxes + b;
          const int cb = (int)(id % boxes_per_row); const long long rb = id / boxes_per_row;
          tma_store(&tout, smem_u32(base + (st * boxes + b) * slot), cb * (WB / 2), (int)(rb * 49));
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      if (s + depth < nstage_total) issue(s + depth);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// store only: each CTA's single thread streams boxes of 49 rows x WB bytes from (uninitialised) shared memory, `inflight` bulk groups deep
template <int WB>
__global__ void store_probe(const __grid_constant__ CUtensorMap tout, int boxes_per_group, int ngroups, int cols_per_row) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  if (threadIdx.x == 0) {
    const int boxes_per_row = cols_per_row / (WB / 2);
    for (int s = 0; s < ngroups; ++s) {
      for (int b = 0; b < boxes_per_group; ++b) {
        const long long id = ((long long)s * gridDim.x + blockIdx.x) * boxes_per_group + b;
        const int cb = (int)(id % boxes_per_row); const long long rb = id / boxes_per_row;
        tma_store(&tout, smem_u32(base + (b & 3) * 64 * WB), cb * (WB / 2), (int)(rb * 49));
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  enc_fn enc = (enc_fn)fp;
  const int cols = 1152;                                  // bf16 columns per row (stage 2: 3C = 1152); row pitch 2304 B
  const long long rows = 49LL * 16384;                    // 1.85 GB
  void *a, *b; cudaMalloc(&a, rows * cols * 2); cudaMalloc(&b, rows * cols * 2);
  cudaMemset(a, 1, rows * cols * 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int wb : {64, 128}) {
    CUtensorMap tin, tout;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)(wb / 2), 49}; cuuint32_t es[2] = {1, 1};
    CUtensorMapSwizzle sw = wb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    enc(&tin, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int cps : {1, 2, 4}) {
      const int boxes_per_group = 2;
      const long long total_boxes = rows / 49 * (cols / (wb / 2));
      const int grid = 148 * cps;
      const int ngroups = (int)(total_boxes / ((long long)grid * boxes_per_group));
      auto k = wb == 64 ? store_probe<64> : store_probe<128>;
      const size_t smem = 4 * 64 * wb + 1024;
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<<<grid, 32, smem>>>(tout, boxes_per_group, ngroups, cols);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("row %3d B  store only       CTAs/SM %d: %6.0f GB/s   [%s]\n", wb, cps, (double)ngroups * grid * boxes_per_group * 49 * wb / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    for (int store_every : {0})
      for (int cps : {1, 2, 4})
        for (int boxes : {1, 2, 4}) {
          const int depth = 4;
          const size_t smem = (size_t)depth * boxes * 64 * wb + 1024;
          if (smem * cps > 220 * 1024) continue;
          const int grid = 148 * cps;
          const long long total_boxes = rows / 49 * (cols / (wb / 2));
          const int nstage = (int)(total_boxes / ((long long)grid * boxes));
          auto k = wb == 64 ? probe<64> : probe<128>;
          cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          float best = 1e30f;
          for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k<<<grid, 32, smem>>>(tin, tout, depth, boxes, nstage, cols, store_every);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
          }
          const double bytes = (double)nstage * grid * boxes * 49 * wb * (store_every ? 1.5 : 1.0);
          printf("row %3d B  %s  CTAs/SM %d  in flight/SM %5.1f KB (depth %d x %2d boxes): %6.0f GB/s   [%s]\n", wb, store_every ? "load+store(1:2)" : "load only      ",
                 cps, (double)cps * depth * boxes * 49 * wb / 1024, depth, boxes, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
  }
  return 0;
}
