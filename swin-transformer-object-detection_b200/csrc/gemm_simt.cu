// fp32 FFMA GEMM (the <=1e-4 parity mode and the on-device cross-check for the tcgen05 path).
// 64x64x16 tiles, 256 threads, 4x4 micro-tiles, optional split-K (ATOMIC_ADD epilogue only).
#include "epilogue.cuh"

namespace swin {

constexpr int SBM = 64, SBN = 64, SBK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, long long sam, long long sak,
                                                        const float* __restrict__ B, long long sbn, long long sbk,
                                                        int K, int k_per_split, EpiParams p) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;      // M tiles on gridDim.x (2^31 limit): M = B*H*W rows can exceed 65535 * 64
  const int k0 = blockIdx.z * k_per_split;
  const int k1 = min(K, k0 + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int kb = k0; kb < k1; kb += SBK) {
    // 64x16 elements each for A and B: 4 per thread.  Map so the contiguous dim is fastest.
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = threadIdx.x + e * 256;
      int mm, kk;
      if (sak == 1) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      int gm = m0 + mm, gk = kb + kk;
      As[kk][mm] = (gm < p.M && gk < k1) ? A[gm * sam + gk * sak] : 0.f;
      int nn;
      if (sbk == 1) { kk = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kk = idx >> 6; }
      int gn = n0 + nn; gk = kb + kk;
      Bs[kk][nn] = (gn < p.N && gk < k1) ? B[gn * sbn + gk * sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = m0 + ty * 4 + i;
    long long drow; float scale;
    if (epi_row_setup(p, row, &drow, &scale)) epilogue_cols<4>(p, row, drow, scale, n0 + tx * 4, acc[i]);
  }
}

// colsum_a for the fp32 path: one thread per m, strided loop over k (A stored (K,M): coalesced across m)
__global__ void simt_colsum_a_kernel(const float* __restrict__ A, long long lda, int M, int K, int k_chunk, float* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int k0 = blockIdx.y * k_chunk, k1 = min(K, k0 + k_chunk);
  float s = 0.f;
  for (int k = k0; k < k1; ++k) s += A[(long long)k * lda + m];
  atomicAdd(out + m, s);
}

int gemm_simt(const swin_gemm_args* a, cudaStream_t st) {
  EpiParams p;
  int rc = make_epi_params(a, &p);
  if (rc) return rc;
  SWIN_REQUIRE(a->A && a->B, "gemm: null operand");
  if (a->M == 0) return 0;
  long long sam = a->a_trans ? 1 : a->lda, sak = a->a_trans ? a->lda : 1;
  long long sbn = a->b_trans ? 1 : a->ldb, sbk = a->b_trans ? a->ldb : 1;
  int gx = ceil_div(a->N, SBN), gy = ceil_div(a->M, SBM);
  int splits = 1;
  if (a->epilogue == SWIN_EPI_ATOMIC_ADD) {
    long long tiles = (long long)gx * gy;
    splits = (int)((4LL * kNumSMs + tiles - 1) / tiles);
    int maxs = ceil_div(a->K, 4 * SBK);
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  int kps = ceil_div(ceil_div(a->K, splits), SBK) * SBK;
  splits = ceil_div(a->K, kps);
  dim3 grid(gy, gx, splits);
  gemm_simt_kernel<<<grid, 256, 0, st>>>((const float*)a->A, sam, sak, (const float*)a->B, sbn, sbk, a->K, kps, p);
  SWIN_LAUNCH_CHECK();
  if (a->colsum_a != nullptr) {
    SWIN_REQUIRE(a->epilogue == SWIN_EPI_ATOMIC_ADD && a->a_trans, "gemm: colsum_a needs ATOMIC_ADD and a_trans=1");
    const int kc = ceil_div(a->K, 4 * kNumSMs) < 64 ? 64 : ceil_div(a->K, 4 * kNumSMs);
    dim3 g2(ceil_div(a->M, 128), ceil_div(a->K, kc));
    simt_colsum_a_kernel<<<g2, 128, 0, st>>>((const float*)a->A, a->lda, a->M, a->K, kc, a->colsum_a);
    SWIN_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace swin
