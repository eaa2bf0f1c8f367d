// Fused window attention on tcgen05 (bf16 operands, fp32 accumulation in TMEM), window 7.
//
// Work item = (window pair, head).  The two 49-token windows are padded to 64 rows each and
// stacked into one 128-row tile (TMEM lane = row):
//     S  = Q K^T        one 128x128x32 UMMA; only the two diagonal 64x64 blocks are used
//     P  = softmax(scale*S + bias + mask)  in registers (thread = row), written to smem as bf16
//     O  = P V          128x32x128 UMMA with P block-diagonal (off-diagonal blocks stay zero)
// Q/K/V tiles arrive by TMA boxes of exactly 49 rows x 32 columns (64-byte swizzle), so the
// bytes moved are the algorithmic ones; pad rows of the tiles are zeroed once and never written.
// Backward recomputes S, forms dP = dO V^T, dS = P*(dP - D), and runs dV = P^T dO, dK = dS^T Q,
// dQ = dS K with MN-major ("transposed") smem descriptors on the same P/dS tiles.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace swin {

constexpr int AHD = 32;          // head dim
constexpr int AN = 49;           // tokens per window (ws = 7)
constexpr uint32_t kTileBytes = 128 * 64;     // one 128-row x 32-col bf16 operand tile (SW64)
constexpr uint32_t kBoxBytes = AN * 64;       // one TMA box
constexpr uint32_t kPBytes = 128 * 128 * 2;   // P / dS tile (two SW128 K-atoms of 16 KB)

struct AttnTcParams {
  int B_, nH, nW, C, npairs, ctas_per_head;
  float scale;
  const float* bias; const float* mask;
  __nv_bfloat16* out; float* lse;
  const __nv_bfloat16* o_saved; const __nv_bfloat16* dout; __nv_bfloat16* dqkv; float* dbias;
};

// byte offset of (row r, 16-byte chunk c) inside a 128-row tile with 128-byte rows, 128B swizzle
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void zero_smem(uint8_t* base, uint32_t bytes) {
  for (uint32_t o = threadIdx.x * 16; o < bytes; o += blockDim.x * 16) *reinterpret_cast<int4*>(base + o) = make_int4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(128, 3) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_slot[2];
  uint8_t* sbase = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = sbase;
  uint8_t* sK = sQ + kTileBytes;
  uint8_t* sV = sK + kTileBytes;
  uint8_t* sP = sV + kTileBytes;                         // 32 KB
  float* sBias = reinterpret_cast<float*>(sP + kPBytes); // [49][49]
  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x % p.nH, g = blockIdx.x / p.nH;
  const uint32_t bar_load = smem_u32(&bars[0]), bar_s = smem_u32(&bars[1]), bar_o = smem_u32(&bars[2]);

  zero_smem(sQ, 3 * kTileBytes + kPBytes);
  for (int e = tid; e < AN * AN; e += blockDim.x) sBias[e] = p.bias[(size_t)h * AN * AN + e];
  if (tid == 0) {
    mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot[0]), 128);
    tmem_alloc(smem_u32(&tmem_slot[1]), 32);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tS = tmem_slot[0], tO = tmem_slot[1];
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  const int r = tid, wloc = r >> 6, i = r & 63;
  const uint32_t idesc_s = umma_idesc_bf16(128, false, false);
  const uint32_t idesc_o = umma_idesc_bf16(32, false, true);
  const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
  const float kLog2e = 1.4426950408889634f;

  uint32_t it = 0;
  for (int pair = g; pair < p.npairs; pair += p.ctas_per_head, ++it) {
    const uint32_t ph = it & 1;
    if (tid == 0) {
      mbar_expect_tx(bar_load, 6 * kBoxBytes);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int row0 = (2 * pair + w) * AN;
        tma_load_2d(aQ + w * 4096, &tmQKV, bar_load, h * AHD, row0);
        tma_load_2d(aK + w * 4096, &tmQKV, bar_load, p.C + h * AHD, row0);
        tma_load_2d(aV + w * 4096, &tmQKV, bar_load, 2 * p.C + h * AHD, row0);
      }
      mbar_wait(bar_load, ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 2; ++k)
        umma_bf16(tS, umma_desc(aQ + k * 32, 16, 512, kSw64), umma_desc(aK + k * 32, 16, 512, kSw64), idesc_s, k);
      umma_commit(bar_s);
    }
    const int win = 2 * pair + wloc;
    const bool valid = (i < AN) && (win < p.B_);
    mbar_wait(bar_s, ph);
    tc_fence_after();
    uint32_t v[64];
    tmem_ld32(tS + lane_off + wloc * 64, v);
    tmem_ld32(tS + lane_off + wloc * 64 + 32, v + 32);
    tmem_ld_wait();
    float mx = -INFINITY, sum = 0.f;
    if (valid) {
      const float* brow = sBias + i * AN;
      const float* mrow = p.mask ? p.mask + ((size_t)(win % p.nW) * AN + i) * AN : nullptr;
#pragma unroll
      for (int j = 0; j < AN; ++j) {
        float s = __uint_as_float(v[j]) * p.scale + brow[j];
        if (mrow) s += __ldg(mrow + j);
        s *= kLog2e;
        v[j] = __float_as_uint(s);
        mx = fmaxf(mx, s);
      }
#pragma unroll
      for (int j = 0; j < AN; ++j) {
        float e = exp2f(__uint_as_float(v[j]) - mx);
        sum += e;
        v[j] = __float_as_uint(e);
      }
#pragma unroll
      for (int j = AN; j < 64; ++j) v[j] = 0u;
    } else {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = 0u;
    }
    // P row -> smem (K-atom `wloc`, 128-byte swizzle)
    {
      uint8_t* prow = sP + wloc * 16384;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        int4 pk;
        pk.x = pack_bf16(__uint_as_float(v[8 * c + 0]), __uint_as_float(v[8 * c + 1]));
        pk.y = pack_bf16(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3]));
        pk.z = pack_bf16(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5]));
        pk.w = pack_bf16(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7]));
        *reinterpret_cast<int4*>(prow + sw128_off(r, c)) = pk;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_bf16(tO, umma_desc(aP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, kSw128),
                  umma_desc(aV + kk * 1024, 512, 512, kSw64), idesc_o, kk);
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld32(tO + lane_off, o);
    tmem_ld_wait();
    if (valid) {
      const float inv = 1.0f / sum;
      __nv_bfloat16* orow = p.out + ((size_t)win * AN + i) * p.C + h * AHD;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int4 pk;
        pk.x = pack_bf16(__uint_as_float(o[8 * c + 0]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
        pk.y = pack_bf16(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
        pk.z = pack_bf16(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
        pk.w = pack_bf16(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
        reinterpret_cast<int4*>(orow)[c] = pk;
      }
      p.lse[((size_t)win * p.nH + h) * AN + i] = (mx + log2f(sum)) * 0.6931471805599453f;   // natural-log LSE
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tS, 128); tmem_dealloc(tO, 32); }
}

// ------------------------------------------------------------------------------------------ backward
__global__ void __launch_bounds__(128, 2) attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                              const __grid_constant__ CUtensorMap tmDO, AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_slot;
  uint8_t* sbase = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = sbase;
  uint8_t* sK = sQ + kTileBytes;
  uint8_t* sV = sK + kTileBytes;
  uint8_t* sdO = sV + kTileBytes;
  uint8_t* sP = sdO + kTileBytes;
  uint8_t* sdS = sP + kPBytes;
  float* sBias = reinterpret_cast<float*>(sdS + kPBytes);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x % p.nH, g = blockIdx.x / p.nH;
  const uint32_t bar_load = smem_u32(&bars[0]), bar_s = smem_u32(&bars[1]), bar_o = smem_u32(&bars[2]);

  zero_smem(sQ, 4 * kTileBytes + 2 * kPBytes);
  for (int e = tid; e < AN * AN; e += blockDim.x) sBias[e] = p.bias[(size_t)h * AN * AN + e];
  if (tid == 0) {
    mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tS = tmem_slot, tdP = tmem_slot + 128;
  const uint32_t tdV = tmem_slot, tdK = tmem_slot + 32, tdQ = tmem_slot + 64;   // reuse S columns after softmax
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  const int r = tid, wloc = r >> 6, i = r & 63;
  const uint32_t idesc_s = umma_idesc_bf16(128, false, false);
  const uint32_t idesc_tt = umma_idesc_bf16(32, true, true);     // A^T B with both MN-major
  const uint32_t idesc_nt = umma_idesc_bf16(32, false, true);
  const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), adO = smem_u32(sdO);
  const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
  const float kLog2e = 1.4426950408889634f;

  float db[AN];
#pragma unroll
  for (int j = 0; j < AN; ++j) db[j] = 0.f;

  uint32_t it = 0;
  for (int pair = g; pair < p.npairs; pair += p.ctas_per_head, ++it) {
    const uint32_t ph = it & 1;
    if (tid == 0) {
      mbar_expect_tx(bar_load, 8 * kBoxBytes);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int row0 = (2 * pair + w) * AN;
        tma_load_2d(aQ + w * 4096, &tmQKV, bar_load, h * AHD, row0);
        tma_load_2d(aK + w * 4096, &tmQKV, bar_load, p.C + h * AHD, row0);
        tma_load_2d(aV + w * 4096, &tmQKV, bar_load, 2 * p.C + h * AHD, row0);
        tma_load_2d(adO + w * 4096, &tmDO, bar_load, h * AHD, row0);
      }
      mbar_wait(bar_load, ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 2; ++k)
        umma_bf16(tS, umma_desc(aQ + k * 32, 16, 512, kSw64), umma_desc(aK + k * 32, 16, 512, kSw64), idesc_s, k);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        umma_bf16(tdP, umma_desc(adO + k * 32, 16, 512, kSw64), umma_desc(aV + k * 32, 16, 512, kSw64), idesc_s, k);
      umma_commit(bar_s);
    }
    const int win = 2 * pair + wloc;
    const bool valid = (i < AN) && (win < p.B_);
    // D_i = <dO_i, O_i> for this head, straight from global while the MMAs run
    float delta = 0.f, lse2 = 0.f;
    if (valid) {
      const int4* orow = reinterpret_cast<const int4*>(p.o_saved + ((size_t)win * AN + i) * p.C + h * AHD);
      const int4* drow = reinterpret_cast<const int4*>(p.dout + ((size_t)win * AN + i) * p.C + h * AHD);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int4 a = __ldg(orow + c), b = __ldg(drow + c);
        const uint32_t au[4] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w};
        const uint32_t bu[4] = {(uint32_t)b.x, (uint32_t)b.y, (uint32_t)b.z, (uint32_t)b.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) delta += bf16_lo(au[e]) * bf16_lo(bu[e]) + bf16_hi(au[e]) * bf16_hi(bu[e]);
      }
      lse2 = p.lse[((size_t)win * p.nH + h) * AN + i] * kLog2e;
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    uint8_t* prow = sP + wloc * 16384;
    uint8_t* dsrow = sdS + wloc * 16384;
    const float* brow = sBias + (valid ? i : 0) * AN;
    const float* mrow = (p.mask && valid) ? p.mask + ((size_t)(win % p.nW) * AN + i) * AN : nullptr;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t s[32], dp[32];
      tmem_ld32(tS + lane_off + wloc * 64 + half * 32, s);
      tmem_ld32(tdP + lane_off + wloc * 64 + half * 32, dp);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const int j = half * 32 + jj;
        float pv = 0.f, ds = 0.f;
        if (valid && j < AN) {
          float sv = __uint_as_float(s[jj]) * p.scale + brow[j];
          if (mrow) sv += __ldg(mrow + j);
          pv = exp2f(sv * kLog2e - lse2);
          ds = pv * (__uint_as_float(dp[jj]) - delta);
          db[j < AN ? j : 0] += ds;
        }
        s[jj] = __float_as_uint(pv);
        dp[jj] = __float_as_uint(ds * p.scale);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int4 pk, dk;
        pk.x = pack_bf16(__uint_as_float(s[8 * c + 0]), __uint_as_float(s[8 * c + 1]));
        pk.y = pack_bf16(__uint_as_float(s[8 * c + 2]), __uint_as_float(s[8 * c + 3]));
        pk.z = pack_bf16(__uint_as_float(s[8 * c + 4]), __uint_as_float(s[8 * c + 5]));
        pk.w = pack_bf16(__uint_as_float(s[8 * c + 6]), __uint_as_float(s[8 * c + 7]));
        dk.x = pack_bf16(__uint_as_float(dp[8 * c + 0]), __uint_as_float(dp[8 * c + 1]));
        dk.y = pack_bf16(__uint_as_float(dp[8 * c + 2]), __uint_as_float(dp[8 * c + 3]));
        dk.z = pack_bf16(__uint_as_float(dp[8 * c + 4]), __uint_as_float(dp[8 * c + 5]));
        dk.w = pack_bf16(__uint_as_float(dp[8 * c + 6]), __uint_as_float(dp[8 * c + 7]));
        *reinterpret_cast<int4*>(prow + sw128_off(r, half * 4 + c)) = pk;
        *reinterpret_cast<int4*>(dsrow + sw128_off(r, half * 4 + c)) = dk;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {   // dV = P^T dO   (K = query rows)
        umma_bf16(tdV, umma_desc(aP + kk * 2048, 16384, 1024, kSw128), umma_desc(adO + kk * 1024, 512, 512, kSw64), idesc_tt, kk);
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {   // dK = (scale dS)^T Q
        umma_bf16(tdK, umma_desc(adS + kk * 2048, 16384, 1024, kSw128), umma_desc(aQ + kk * 1024, 512, 512, kSw64), idesc_tt, kk);
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {   // dQ = (scale dS) K   (K = key rows)
        umma_bf16(tdQ, umma_desc(adS + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, kSw128),
                  umma_desc(aK + kk * 1024, 512, 512, kSw64), idesc_nt, kk);
      }
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
#pragma unroll
    for (int part = 0; part < 3; ++part) {   // 0: dQ, 1: dK, 2: dV  (column blocks of dqkv)
      uint32_t o[32];
      tmem_ld32((part == 0 ? tdQ : part == 1 ? tdK : tdV) + lane_off, o);
      tmem_ld_wait();
      if (valid) {
        __nv_bfloat16* orow = p.dqkv + ((size_t)win * AN + i) * 3 * p.C + part * p.C + h * AHD;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          int4 pk;
          pk.x = pack_bf16(__uint_as_float(o[8 * c + 0]), __uint_as_float(o[8 * c + 1]));
          pk.y = pack_bf16(__uint_as_float(o[8 * c + 2]), __uint_as_float(o[8 * c + 3]));
          pk.z = pack_bf16(__uint_as_float(o[8 * c + 4]), __uint_as_float(o[8 * c + 5]));
          pk.w = pack_bf16(__uint_as_float(o[8 * c + 6]), __uint_as_float(o[8 * c + 7]));
          reinterpret_cast<int4*>(orow)[c] = pk;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (i < AN) {
    float* dst = p.dbias + ((size_t)h * AN + i) * AN;
#pragma unroll
    for (int j = 0; j < AN; ++j) atomicAdd(dst + j, db[j]);
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_slot, 256); }
}

static int attn_tc_common(const swin_attn_args* a, AttnTcParams* out, bool bwd, int ctas_per_sm) {
  SWIN_REQUIRE(a->ws == 7, "attn(bf16): the tcgen05 kernel supports window_size 7 (got %d)", a->ws);
  SWIN_REQUIRE(a->B_ >= 0 && a->nH > 0, "attn: bad shape");
  SWIN_REQUIRE(a->qkv && a->bias && a->lse && a->out, "attn: null pointer");
  SWIN_REQUIRE(a->mask == nullptr || (a->nW > 0 && a->B_ % a->nW == 0), "attn: B_ must be a multiple of nW when a mask is given");
  SWIN_REQUIRE(aligned16(a->qkv) && aligned16(a->out), "attn: alignment");
  if (bwd) SWIN_REQUIRE(a->dout && a->dqkv && a->dbias && aligned16(a->dout) && aligned16(a->dqkv), "attn_bwd: null/misaligned pointer");
  AttnTcParams p;
  p.B_ = a->B_; p.nH = a->nH; p.nW = a->nW > 0 ? a->nW : 1; p.C = a->nH * AHD; p.scale = a->scale;
  p.npairs = (a->B_ + 1) / 2;
  int per_head = ceil_div(kNumSMs * ctas_per_sm, a->nH);
  if (per_head > p.npairs) per_head = p.npairs;
  if (per_head < 1) per_head = 1;
  p.ctas_per_head = per_head;
  p.bias = a->bias; p.mask = a->mask; p.out = (__nv_bfloat16*)a->out; p.lse = a->lse;
  p.o_saved = (const __nv_bfloat16*)a->out; p.dout = (const __nv_bfloat16*)a->dout; p.dqkv = (__nv_bfloat16*)a->dqkv; p.dbias = a->dbias;
  *out = p;
  return 0;
}

int attn_tc_fwd(const swin_attn_args* a, cudaStream_t st) {
  AttnTcParams p;
  int rc = attn_tc_common(a, &p, false, 3);
  if (rc) return rc;
  if (p.B_ == 0) return 0;
  CUtensorMap tm;
  rc = make_tmap_bf16_2d(&tm, a->qkv, (uint64_t)3 * p.C, (uint64_t)p.B_ * AN, (uint64_t)3 * p.C * 2, AHD, AN, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  const size_t smem = 3 * kTileBytes + kPBytes + AN * AN * sizeof(float) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  attn_tc_fwd_kernel<<<p.nH * p.ctas_per_head, 128, smem, st>>>(tm, p);
  SWIN_LAUNCH_CHECK();
  return 0;
}

int attn_tc_bwd(const swin_attn_args* a, cudaStream_t st) {
  AttnTcParams p;
  int rc = attn_tc_common(a, &p, true, 2);
  if (rc) return rc;
  if (p.B_ == 0) return 0;
  CUtensorMap tm, tmdo;
  rc = make_tmap_bf16_2d(&tm, a->qkv, (uint64_t)3 * p.C, (uint64_t)p.B_ * AN, (uint64_t)3 * p.C * 2, AHD, AN, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmdo, a->dout, (uint64_t)p.C, (uint64_t)p.B_ * AN, (uint64_t)p.C * 2, AHD, AN, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  const size_t smem = 4 * kTileBytes + 2 * kPBytes + AN * AN * sizeof(float) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  attn_tc_bwd_kernel<<<p.nH * p.ctas_per_head, 128, smem, st>>>(tm, tmdo, p);
  SWIN_LAUNCH_CHECK();
  return 0;
}

}  // namespace swin
