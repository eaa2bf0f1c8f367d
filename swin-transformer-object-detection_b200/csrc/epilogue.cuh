// Fused GEMM epilogues shared by the FFMA (fp32) and tcgen05 (bf16) GEMM kernels.
// One call handles NV consecutive columns of one accumulator row.
#pragma once
#include "common.cuh"

namespace swin {

struct EpiParams {
  int M, N;
  int epilogue;
  const float* bias;
  void* D;
  int d_dtype;
  long long ldd;
  void* D2;
  const void* aux;
  const float* row_scale;
  int rows_per_image;
  WinGeom g;   // SCATTER_RESIDUAL only
};

__device__ __forceinline__ void store4(void* base, int dtype, long long idx, float4 v) {
  if (dtype == SWIN_F32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = v;
  else *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
}
__device__ __forceinline__ float4 load4(const void* base, int dtype, long long idx) {
  if (dtype == SWIN_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
  uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
  return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
}

// Resolve the destination row and drop-path scale of accumulator row `row`.
// Returns false when the row produces no output (out of range, or a padded window slot).
__device__ __forceinline__ bool epi_row_setup(const EpiParams& p, int row, long long* drow, float* scale) {
  if (row >= p.M) return false;
  *drow = row;
  *scale = 1.0f;
  if (p.epilogue == SWIN_EPI_RESIDUAL) {
    if (p.row_scale) *scale = p.row_scale[row / p.rows_per_image];
  } else if (p.epilogue == SWIN_EPI_SCATTER_RESIDUAL) {
    int b = row / p.rows_per_image;
    int tok = slot_to_token(p.g, row - b * p.rows_per_image);
    if (tok < 0) return false;
    *drow = (long long)b * (p.g.H * p.g.W) + tok;
    if (p.row_scale) *scale = p.row_scale[b];
  }
  return true;
}

template <int NV, bool FAST = false>
__device__ __forceinline__ void epilogue_cols(const EpiParams& p, int row, long long drow, float scale, int col0, const float* v) {
#pragma unroll
  for (int c = 0; c < NV; c += 4) {
    const int col = col0 + c;
    if (col >= p.N) break;
    float4 a = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
    if (p.bias != nullptr) {
      float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
    }
    const long long o = drow * p.ldd + col;
    switch (p.epilogue) {
      case SWIN_EPI_STORE:
        store4(p.D, p.d_dtype, o, a);
        break;
      case SWIN_EPI_GELU:                       // D = gelu(u), D2 = gelu'(u)  (u = acc + bias)
        if (FAST) {
          float4 gq, dq;
          gelu_fast(a.x, &gq.x, &dq.x); gelu_fast(a.y, &gq.y, &dq.y); gelu_fast(a.z, &gq.z, &dq.z); gelu_fast(a.w, &gq.w, &dq.w);
          store4(p.D, p.d_dtype, o, gq);
          store4(p.D2, p.d_dtype, o, dq);
        } else {
          store4(p.D, p.d_dtype, o, make_float4(gelu_erf(a.x), gelu_erf(a.y), gelu_erf(a.z), gelu_erf(a.w)));
          store4(p.D2, p.d_dtype, o, make_float4(dgelu_erf(a.x), dgelu_erf(a.y), dgelu_erf(a.z), dgelu_erf(a.w)));
        }
        break;
      case SWIN_EPI_RESIDUAL:
      case SWIN_EPI_SCATTER_RESIDUAL: {
        float4 r = load4(p.aux, SWIN_F32, o);
        store4(p.D, SWIN_F32, o, make_float4(r.x + scale * a.x, r.y + scale * a.y, r.z + scale * a.z, r.w + scale * a.w));
        break;
      }
      case SWIN_EPI_DGELU: {                    // D = acc * aux  (aux = gelu'(u) saved by the GELU epilogue)
        float4 g = load4(p.aux, p.d_dtype, o);
        store4(p.D, p.d_dtype, o, make_float4(a.x * g.x, a.y * g.y, a.z * g.z, a.w * g.w));
        break;
      }
      case SWIN_EPI_ATOMIC_ADD: {
        float* d = reinterpret_cast<float*>(p.D) + o;
        atomicAdd(d + 0, a.x); atomicAdd(d + 1, a.y); atomicAdd(d + 2, a.z); atomicAdd(d + 3, a.w);
        break;
      }
      default:
        break;
    }
  }
}

// Host-side validation + packing shared by both GEMM front-ends.
inline int make_epi_params(const swin_gemm_args* a, EpiParams* out) {
  EpiParams p;
  SWIN_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, "gemm: bad M/N/K");
  SWIN_REQUIRE(a->N % 8 == 0, "gemm: N must be a multiple of 8");
  SWIN_REQUIRE(a->ldd % 8 == 0 && a->ldd >= a->N, "gemm: ldd must be >= N and a multiple of 8");
  SWIN_REQUIRE(a->D != nullptr && aligned16(a->D), "gemm: D null or misaligned");
  SWIN_REQUIRE(a->d_dtype == SWIN_F32 || a->d_dtype == SWIN_BF16, "gemm: bad d_dtype");
  SWIN_REQUIRE(a->bias == nullptr || aligned16(a->bias), "gemm: bias misaligned");
  p.M = a->M; p.N = a->N; p.epilogue = a->epilogue; p.bias = a->bias; p.D = a->D; p.d_dtype = a->d_dtype;
  p.ldd = a->ldd; p.D2 = a->D2; p.aux = a->aux; p.row_scale = a->row_scale; p.rows_per_image = a->rows_per_image;
  p.g = make_geom(1, 1, 1, 1, 1, 0);
  switch (a->epilogue) {
    case SWIN_EPI_STORE: break;
    case SWIN_EPI_GELU:
      SWIN_REQUIRE(a->D2 != nullptr && aligned16(a->D2), "gemm: GELU epilogue needs D2");
      break;
    case SWIN_EPI_RESIDUAL:
      SWIN_REQUIRE(a->aux != nullptr && aligned16(a->aux) && a->d_dtype == SWIN_F32, "gemm: RESIDUAL needs fp32 aux/D");
      SWIN_REQUIRE(a->row_scale == nullptr || a->rows_per_image > 0, "gemm: rows_per_image");
      if (p.rows_per_image <= 0) p.rows_per_image = a->M > 0 ? a->M : 1;
      break;
    case SWIN_EPI_SCATTER_RESIDUAL: {
      SWIN_REQUIRE(a->aux != nullptr && aligned16(a->aux) && a->d_dtype == SWIN_F32, "gemm: SCATTER_RESIDUAL needs fp32 aux/D");
      SWIN_REQUIRE(a->H > 0 && a->W > 0 && a->ws > 0 && a->shift >= 0 && a->shift < a->ws, "gemm: bad scatter geometry");
      p.g = make_geom(1, a->H, a->W, a->N, a->ws, a->shift);
      p.rows_per_image = p.g.nW * p.g.N;
      SWIN_REQUIRE(a->M % p.rows_per_image == 0, "gemm: M is not a whole number of images' window slots");
      break;
    }
    case SWIN_EPI_DGELU:
      SWIN_REQUIRE(a->aux != nullptr && aligned16(a->aux), "gemm: DGELU needs aux = saved gelu'");
      break;
    case SWIN_EPI_ATOMIC_ADD:
      SWIN_REQUIRE(a->d_dtype == SWIN_F32 && a->bias == nullptr, "gemm: ATOMIC_ADD needs fp32 D and no bias");
      break;
    default:
      set_error("gemm: unknown epilogue %d", a->epilogue);
      return -EINVAL;
  }
  *out = p;
  return 0;
}

}  // namespace swin
