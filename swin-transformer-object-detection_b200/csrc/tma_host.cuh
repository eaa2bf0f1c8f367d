// Host-side TMA descriptor construction (cuTensorMapEncodeTiled through the runtime's
// driver-entry-point query, so the library links against cudart only).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace swin {

typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline tmap_encode_fn get_tmap_encode() {
  static tmap_encode_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (tmap_encode_fn)p;
  }();
  return fn;
}

// 2-D bf16 tensor, row-major: `inner` contiguous elements per row, `outer` rows, row pitch in bytes.
// Out-of-bounds box elements are filled with zeros.
inline int make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                        uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz);
inline int make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                             uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  return make_tmap_2d(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ptr, inner, outer, pitch_bytes, box_inner, box_outer, swz);
}
inline int make_tmap_f32_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                            uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  return make_tmap_2d(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ptr, inner, outer, pitch_bytes, box_inner, box_outer, swz);
}
inline int make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                        uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  tmap_encode_fn enc = get_tmap_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return -ENOTSUP; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (pitch_bytes & 15)) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch (ptr=%p pitch=%llu)", ptr,
              (unsigned long long)pitch_bytes);
    return -EINVAL;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims=%llu x %llu pitch=%llu box=%u x %u", (int)r, (unsigned long long)inner,
              (unsigned long long)outer, (unsigned long long)pitch_bytes, box_inner, box_outer);
    return -EINVAL;
  }
  return 0;
}

}  // namespace swin
