// Fused window attention on tcgen05 (bf16 operands, fp32 accumulation in TMEM), window 7.
//
// Work item = (window pair, head).  The two 49-token windows are padded to 64 rows each and
// stacked into one 128-row tile (TMEM lane = row (w, i)):
//     S  = Q K^T     one 128x128x32 UMMA; only the two diagonal 64x64 blocks are used
//     P  = softmax(scale*S + bias + mask)  in registers (thread = row), written to smem as a
//          COMPACT 128x64 bf16 tile (row (w,i), column j)
//     O  = P [V_0 | V_1]   128x64x64 UMMA: the B operand is both windows' V tiles side by side
//          (N = (w', d)); row (w,i) keeps columns w'=w.
// Q/K/V tiles arrive by TMA boxes of exactly 49 rows x 32 columns (64-byte swizzle) so the bytes
// moved are the algorithmic ones; pad rows of the tiles are zeroed once and never written.
// Forward: 128-thread CTAs (thread = row), four per SM, single-buffered tiles refilled as soon as the MMA that read
// them has completed.  Backward: 256-thread CTAs (two threads per row), two per SM, double-buffered tiles.
// Backward recomputes S, forms dP = dO V^T, D = rowsum(P*dP), dS = P*(dP - D), and runs
//     dV = P^T [dO_0|dO_1],  dK = dS^T [Q_0|Q_1],  dQ = dS [K_0|K_1]
// with MN-major ("transposed") smem descriptors over the same compact P / dS tiles.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace swin {

constexpr int AHD = 32;          // head dim
constexpr int AN = 49;           // tokens per window (ws = 7)
constexpr uint32_t kTileBytes = 128 * 64;     // one 128-row x 32-col bf16 operand tile (SW64); window 1 at +4096
constexpr uint32_t kBoxBytes = AN * 64;       // one TMA box
constexpr uint32_t kPBytes = 128 * 128;       // compact P / dS tile: 128 rows x 64 bf16 (one SW128 atom wide)

struct AttnTcParams {
  int B_, nH, nW, C, npairs, ctas_per_head;
  int canon_nwh, canon_nww;    // > 0: canonical shift mask in closed form
  float scale;
  const float* bias; const float* mask; const int* mask_nz;
  __nv_bfloat16* out; float* lse;
  const __nv_bfloat16* dout; __nv_bfloat16* dqkv; float* dbias;
};

// byte offset of (row r, 16-byte chunk c) inside a tile with 128-byte rows, 128B swizzle
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void zero_smem(uint8_t* base, uint32_t bytes) {
  for (uint32_t o = threadIdx.x * 16; o < bytes; o += blockDim.x * 16) *reinterpret_cast<int4*>(base + o) = make_int4(0, 0, 0, 0);
}

__device__ __forceinline__ void store_row_bf16x8(uint8_t* dst, const float* v) {
  int4 pk;
  pk.x = pack_bf16(v[0], v[1]); pk.y = pack_bf16(v[2], v[3]); pk.z = pack_bf16(v[4], v[5]); pk.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<int4*>(dst) = pk;
}


// Canonical SW-MSA mask (REF:370-389) of one window row in closed form, window 7 / shift 3.  Region ids on the
// padded grid differ only inside the last window row / column, where tokens with row (col) index >= ws - shift = 4
// belong to another region than those < 4.  Returns a 49-bit set: bit j = 1 <=> mask[w][i][j] == -100.
__device__ __forceinline__ unsigned long long canon_mask_bits(int wi, int nwh, int nww, int i) {
  constexpr unsigned long long kRowHi = 0x1FFFFF0000000ULL;          // j/7 >= 4  <=> j >= 28   (bits 28..48)
  constexpr unsigned long long kColHi = 0x1C3870E1C3870ULL;          // j%7 >= 4  (bits {4,5,6} + 7k, k = 0..6)
  constexpr unsigned long long kAll = 0x1FFFFFFFFFFFFULL;
  const int wh = wi / nww, ww = wi - wh * nww;
  unsigned long long m = 0ULL;
  if (wh == nwh - 1) m |= (i / 7 >= 4) ? (~kRowHi & kAll) : kRowHi;
  if (ww == nww - 1) m |= (i % 7 >= 4) ? (~kColHi & kAll) : kColHi;
  return m;
}

// Thread mapping of the backward kernel: 256 threads = 8 warps.  Warp w owns TMEM lane quarter q = w & 3 (rows
// q*32..q*32+31 of the stacked tile) and column half hf = w >> 2 of the row's own 64-column window block, so two
// threads cooperate on one row (row statistics are exchanged through smem).  This doubles the warps available to
// hide the LDS / TMEM / MUFU latencies of the per-row softmax math.
constexpr int kBiasLd = 68;                      // sBias row pitch (floats): 64 columns + 4 so that 8 consecutive rows' 16-byte
                                                 // reads fall in distinct bank groups; pre-scaled by log2(e); columns >= 49 hold
                                                 // kNegBig so padded keys drop out of the softmax with no per-element predicate
constexpr float kNegBig = -1.0e30f;
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr int kAttnThreads = 256;

__device__ __forceinline__ void load_bias_tile(float* sBias, const float* __restrict__ bias, int h) {
  const float kLog2e = 1.4426950408889634f;
  for (int e = threadIdx.x; e < AN * kBiasLd; e += blockDim.x) {
    const int i = e / kBiasLd, j = e - i * kBiasLd;
    sBias[e] = j < AN ? bias[((size_t)h * AN + i) * AN + j] * kLog2e : kNegBig;
  }
}

// Development aid: -DSWIN_ATTN_TIMING accumulates per-phase clock64() deltas of one softmax thread of CTA 0 and
// prints them at kernel exit (never enabled in the shipped library).
#ifdef SWIN_ATTN_TIMING
#define TDECL long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tprev = clock64(); int titems = 0;
#define TMARK(k) do { if (blockIdx.x == 0 && tid == 40) { long long t_ = clock64(); tacc[k] += t_ - tprev; tprev = t_; } } while (0)
#define TITEM ++titems;
#define TPRINT(name) do { if (blockIdx.x == 0 && tid == 40) printf("%s items=%d cycles/item: load+S %lld | tmem_ld %lld | math1 %lld | sync1 %lld | math2 %lld | Pstore+sync %lld | mmaO wait %lld | epilogue %lld | endsync %lld\n", name, titems, tacc[0] / titems, tacc[1] / titems, tacc[2] / titems, tacc[3] / titems, tacc[4] / titems, tacc[5] / titems, tacc[6] / titems, tacc[7] / titems, tacc[8] / titems); } while (0)
#else
#define TDECL
#define TMARK(k)
#define TITEM
#define TPRINT(name)
#endif

// ------------------------------------------------------------------------------------------ forward
// 128 threads per CTA, thread = one row of the stacked 128-row tile (all 49 columns of its own window), four CTAs per SM:
// an item is a strictly serial chain (TMA -> S MMA -> softmax -> P.V MMA -> O -> TMA store) whose latencies are hidden by
// the other three resident CTAs rather than by intra-CTA pipelining; that needs <= 56 KB of smem (single-buffered tiles:
// Q,K are re-filled as soon as S has been computed, V as soon as P.V has, so the next item's loads still overlap this
// item's softmax), <= 128 registers and 128 TMEM columns (O overlays columns [0,64) of S once P is in smem).
constexpr uint32_t kFwdTiles = 3 * kTileBytes;     // Q,K,V
constexpr int kFwdThreads = 128;
constexpr int kFwdBiasLd = 52;                     // bias row pitch of the forward kernel: 49 columns + 3 x kNegBig

__global__ void __launch_bounds__(kFwdThreads, 4) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                      const __grid_constant__ CUtensorMap tmOut, AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4];              // qk, v, s, o
  __shared__ uint32_t tmem_slot;
  uint8_t* sbase = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sT = sbase;                                   // {Q,K,V}
  uint8_t* sP = sT + kFwdTiles;                          // 16 KB: P, later the O staging tile
  float* sBias = reinterpret_cast<float*>(sP + kPBytes); // [49][52]
  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x % p.nH, g = blockIdx.x / p.nH;
  const uint32_t bar_qk = smem_u32(&bars[0]), bar_v = smem_u32(&bars[1]), bar_s = smem_u32(&bars[2]), bar_o = smem_u32(&bars[3]);

  zero_smem(sT, kFwdTiles + kPBytes);
  {
    const float kLog2e = 1.4426950408889634f;
    for (int e = tid; e < AN * kFwdBiasLd; e += blockDim.x) {
      const int bi = e / kFwdBiasLd, bj = e - bi * kFwdBiasLd;
      sBias[e] = bj < AN ? p.bias[((size_t)h * AN + bi) * AN + bj] * kLog2e : kNegBig;
    }
  }
  if (tid == 0) {
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tS = tmem_slot, tO = tmem_slot;
  const int r = tid, wloc = r >> 6, i = r & 63;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t idesc_s = umma_idesc_bf16(128, false, false);
  const uint32_t idesc_o = umma_idesc_bf16(64, false, true);
  const uint32_t aQ = smem_u32(sT), aK = aQ + kTileBytes, aV = aK + kTileBytes, aP = smem_u32(sP);
  const float kLog2e = 1.4426950408889634f;
  const float sc2 = p.scale * kLog2e;
  const int stride = p.ctas_per_head;

  auto issue_qk = [&](int pair) {
    mbar_expect_tx(bar_qk, 4 * kBoxBytes);
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int row0 = (2 * pair + w) * AN;
      tma_load_2d(aQ + w * 4096, &tmQKV, bar_qk, h * AHD, row0);
      tma_load_2d(aK + w * 4096, &tmQKV, bar_qk, p.C + h * AHD, row0);
    }
  };
  auto issue_v = [&](int pair) {
    mbar_expect_tx(bar_v, 2 * kBoxBytes);
#pragma unroll
    for (int w = 0; w < 2; ++w) tma_load_2d(aV + w * 4096, &tmQKV, bar_v, 2 * p.C + h * AHD, (2 * pair + w) * AN);
  };
  // Single-thread work (TMA / MMA issue) sits on the item's serial chain: warp 0 enters converged and ONE elected lane issues
  // (no per-active-lane retry loops in the SASS), with descriptors built as constant-hi : incremented-lo words.
  constexpr uint32_t kHi64 = umma_desc_hi(512, (uint32_t)kSw64), kHi128 = umma_desc_hi(1024, (uint32_t)kSw128);
  const uint32_t q_lo = umma_desc_lo(aQ, 16), k_lo = umma_desc_lo(aK, 16), p_lo = umma_desc_lo(aP, 16), v_lo = umma_desc_lo(aV, 4096);
  if (warp == 0) {
    if (elect_one() && g < p.npairs) { issue_qk(g); issue_v(g); }
    __syncwarp();
  }

  uint32_t it = 0;
  for (int pair = g; pair < p.npairs; pair += stride, ++it) {
    const uint32_t ph = it & 1;
    const bool has_next = pair + stride < p.npairs;
    if (warp == 0) {                             // S = Q K^T as soon as Q, K have landed
      if (elect_one()) {
        mbar_wait(bar_qk, ph);
        tc_fence_after();
#pragma unroll
        for (uint32_t k = 0; k < 2; ++k)
          umma_bf16(tS, umma_desc_join(kHi64, q_lo + 2 * k), umma_desc_join(kHi64, k_lo + 2 * k), idesc_s, k);
        umma_commit(bar_s);
      }
      __syncwarp();
    }
    const int win = 2 * pair + wloc;
    const bool valid = (i < AN) && (win < p.B_);
    const float* mrow = nullptr;
    unsigned long long mb = 0ULL;                // closed-form mask bits of this row (bit j: -100)
    if (p.mask != nullptr && valid) {
      const int mw = win % p.nW;
      if (p.canon_nwh > 0) mb = canon_mask_bits(mw, p.canon_nwh, p.canon_nww, i);
      else if (p.mask_nz == nullptr || p.mask_nz[mw]) mrow = p.mask + ((size_t)mw * AN + i) * AN;
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    if (warp == 0) {                             // the Q, K tiles are free again
      if (elect_one() && has_next) issue_qk(pair + stride);
      __syncwarp();
    }
    // Rows >= 49 of a window (and a whole missing window) run the same math on harmless finite values: their P rows only
    // feed O rows that are never stored.  Columns 49..51 carry kNegBig from the bias tile, so they become exact zeros.
    uint32_t v[52];
    tmem_ld32(tS + lane_off + wloc * 64, v);
    tmem_ld16(tS + lane_off + wloc * 64 + 32, v + 32);
    tmem_ld4(tS + lane_off + wloc * 64 + 48, v + 48);
    tmem_ld_wait();
    float sv[52];
    {
      const float4* b4 = reinterpret_cast<const float4*>(sBias + (i < AN ? i : AN - 1) * kFwdBiasLd);
#pragma unroll
      for (int c = 0; c < 13; ++c) {
        const float4 bb = b4[c];
        sv[4 * c + 0] = fmaf(__uint_as_float(v[4 * c + 0]), sc2, bb.x);
        sv[4 * c + 1] = fmaf(__uint_as_float(v[4 * c + 1]), sc2, bb.y);
        sv[4 * c + 2] = fmaf(__uint_as_float(v[4 * c + 2]), sc2, bb.z);
        sv[4 * c + 3] = fmaf(__uint_as_float(v[4 * c + 3]), sc2, bb.w);
      }
    }
    if (mrow != nullptr) {
#pragma unroll
      for (int jj = 0; jj < AN; ++jj) sv[jj] = fmaf(__ldg(mrow + jj), kLog2e, sv[jj]);
    }
    if (mb != 0ULL) {
      const uint32_t lo = (uint32_t)mb, hi = (uint32_t)(mb >> 32);
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        if ((lo >> jj) & 1u) sv[jj] -= 100.0f * kLog2e;
#pragma unroll
      for (int jj = 32; jj < AN; ++jj)
        if ((hi >> (jj - 32)) & 1u) sv[jj] -= 100.0f * kLog2e;
    }
    float mx = sv[0];
#pragma unroll
    for (int jj = 1; jj < AN; ++jj) mx = fmaxf(mx, sv[jj]);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 52; ++jj) {
      const float e = ex2_ftz(sv[jj] - mx);
      sum += e;
      sv[jj] = e;
    }
    // P row -> compact 128x64 bf16 tile (SW128): chunks 0..5 = columns 0..47, chunk 6 = columns 48..51 + zeros, chunk 7 = zeros
#pragma unroll
    for (int c = 0; c < 6; ++c) store_row_bf16x8(sP + sw128_off(r, c), sv + 8 * c);
    {
      int4 pk;
      pk.x = pack_bf16(sv[48], sv[49]); pk.y = pack_bf16(sv[50], sv[51]); pk.z = 0; pk.w = 0;
      *reinterpret_cast<int4*>(sP + sw128_off(r, 6)) = pk;
      *reinterpret_cast<int4*>(sP + sw128_off(r, 7)) = make_int4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {                             // O = P [V0 | V1]
      if (elect_one()) {
        mbar_wait(bar_v, ph);
        tc_fence_after();
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk)
          umma_bf16(tO, umma_desc_join(kHi128, p_lo + 2 * kk), umma_desc_join(kHi64, v_lo + 64 * kk), idesc_o, kk);
        umma_commit(bar_o);
      }
      __syncwarp();
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if (warp == 0) {                             // the V tile is free again
      if (elect_one() && has_next) issue_v(pair + stride);
      __syncwarp();
    }
    uint32_t o[32];
    tmem_ld32(tO + lane_off + wloc * 32, o);
    tmem_ld_wait();
    {
      // O row -> bf16 tile in sP (free: the P.V MMA has completed), 64-byte-swizzle layout, window w at +4096; then one
      // TMA store of 49 rows x 32 columns per window
      const float inv = valid ? 1.0f / sum : 0.f;
      const uint32_t swz = (uint32_t)((r >> 1) & 3);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = __uint_as_float(o[8 * c + e]) * inv;
        store_row_bf16x8(sP + r * 64 + (((uint32_t)c ^ swz) << 4), t);
      }
      if (valid) p.lse[((size_t)win * p.nH + h) * AN + i] = (mx + log2f(sum)) * 0.6931471805599453f;   // natural-log LSE
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
#pragma unroll
        for (int w = 0; w < 2; ++w)
          if (2 * pair + w < p.B_) tma_store_2d(&tmOut, aP + w * 4096, h * AHD, (2 * pair + w) * AN);
        tma_store_commit();
        // sP is rewritten with the next item's P only after every thread has passed bar_s of the next item, which this
        // thread commits after this wait: the staged O tile has been read out by then
        tma_store_wait_read<0>();
      }
      __syncwarp();
    }
  }
  if (warp == 0) {
    if (elect_one()) tma_store_wait_all<0>();
    __syncwarp();
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_slot, 128); }
}

// ------------------------------------------------------------------------------------------ backward
constexpr uint32_t kBwdTiles = 4 * kTileBytes;     // Q,K,V,dO per buffer

__global__ void __launch_bounds__(kAttnThreads, 2) attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                       const __grid_constant__ CUtensorMap tmDO,
                                                                       const __grid_constant__ CUtensorMap tmDQKV, AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  __shared__ float sRed[2][128];                          // partial D per column half
  uint8_t* sbase = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sT = sbase;                     // 2 x {Q,K,V,dO}
  uint8_t* sP = sT + 2 * kBwdTiles;
  uint8_t* sdS = sP + kPBytes;
  float* sBias = reinterpret_cast<float*>(sdS + kPBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;
  const int h = blockIdx.x % p.nH, g = blockIdx.x / p.nH;
  const uint32_t bar_load0 = smem_u32(&bars[0]), bar_s = smem_u32(&bars[2]), bar_o = smem_u32(&bars[3]);

  zero_smem(sT, 2 * kBwdTiles + 2 * kPBytes);
  load_bias_tile(sBias, p.bias, h);
  if (tid == 0) {
    mbar_init(bar_load0, 1); mbar_init(bar_load0 + 8, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
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
  const uint32_t tdV = tmem_slot, tdK = tmem_slot + 64, tdQ = tmem_slot + 128;   // reuse S / dP columns after the softmax pass
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;

  const int r = q * 32 + lane, wloc = r >> 6, i = r & 63;
  const int jbase = hf * 32;
  const uint32_t idesc_s = umma_idesc_bf16(128, false, false);
  const uint32_t idesc_tt = umma_idesc_bf16(64, true, true);     // A^T B, both MN-major
  const uint32_t idesc_nt = umma_idesc_bf16(64, false, true);
  const uint32_t aT = smem_u32(sT), aP = smem_u32(sP), adS = smem_u32(sdS);
  const float kLog2e = 1.4426950408889634f;
  const float sc2 = p.scale * kLog2e;

  auto issue_loads = [&](int pair, int buf) {
    const uint32_t bar = bar_load0 + 8 * buf, base = aT + buf * kBwdTiles;
    mbar_expect_tx(bar, 8 * kBoxBytes);
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int row0 = (2 * pair + w) * AN;
      tma_load_2d(base + w * 4096, &tmQKV, bar, h * AHD, row0);
      tma_load_2d(base + kTileBytes + w * 4096, &tmQKV, bar, p.C + h * AHD, row0);
      tma_load_2d(base + 2 * kTileBytes + w * 4096, &tmQKV, bar, 2 * p.C + h * AHD, row0);
      tma_load_2d(base + 3 * kTileBytes + w * 4096, &tmDO, bar, h * AHD, row0);
    }
  };
  constexpr uint32_t kHi64 = umma_desc_hi(512, (uint32_t)kSw64), kHi128 = umma_desc_hi(1024, (uint32_t)kSw128);
  const uint32_t p_lo_k = umma_desc_lo(aP, 16), ds_lo_k = umma_desc_lo(adS, 16);              // K-major views of P / dS
  const uint32_t p_lo_mn = umma_desc_lo(aP, 8192), ds_lo_mn = umma_desc_lo(adS, 8192);          // MN-major (transposed) views
  (void)p_lo_k;
  if (warp == 0) {
    if (elect_one() && g < p.npairs) issue_loads(g, 0);
    __syncwarp();
  }

  float db[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) db[j] = 0.f;
  TDECL

  uint32_t it = 0;
  for (int pair = g; pair < p.npairs; pair += p.ctas_per_head, ++it) {
    const uint32_t ph = it & 1, buf = it & 1, lph = (it >> 1) & 1;
    const uint32_t aQ = aT + buf * kBwdTiles, aK = aQ + kTileBytes, aV = aK + kTileBytes, adO = aV + kTileBytes;
    TITEM
    // operand tiles of this buffer as descriptor low words (K-major views: LBO 16; MN-major views of Q, K, dO: LBO 4096)
    const uint32_t q_lo = umma_desc_lo(aQ, 16), k_lo = umma_desc_lo(aK, 16), v_lo = umma_desc_lo(aV, 16), do_lo = umma_desc_lo(adO, 16);
    const uint32_t q_lo_mn = umma_desc_lo(aQ, 4096), k_lo_mn = umma_desc_lo(aK, 4096), do_lo_mn = umma_desc_lo(adO, 4096);
    if (warp == 0) {
      if (elect_one()) {
        mbar_wait(bar_load0 + 8 * buf, lph);
        tc_fence_after();
#pragma unroll
        for (uint32_t k = 0; k < 2; ++k)
          umma_bf16(tS, umma_desc_join(kHi64, q_lo + 2 * k), umma_desc_join(kHi64, k_lo + 2 * k), idesc_s, k);
#pragma unroll
        for (uint32_t k = 0; k < 2; ++k)
          umma_bf16(tdP, umma_desc_join(kHi64, do_lo + 2 * k), umma_desc_join(kHi64, v_lo + 2 * k), idesc_s, k);
        umma_commit(bar_s);
        if (pair + p.ctas_per_head < p.npairs) issue_loads(pair + p.ctas_per_head, buf ^ 1);   // after the MMAs are in flight
      }
      __syncwarp();
    }
    const int win = 2 * pair + wloc;
    const bool valid = (i < AN) && (win < p.B_);
    const float lse2 = valid ? p.lse[((size_t)win * p.nH + h) * AN + i] * kLog2e : 0.f;
    const float* mrow = nullptr;
    uint32_t mb = 0u;
    if (p.mask != nullptr && valid) {
      const int mw = win % p.nW;
      if (p.canon_nwh > 0) mb = (uint32_t)(canon_mask_bits(mw, p.canon_nwh, p.canon_nww, i) >> jbase);
      else if (p.mask_nz == nullptr || p.mask_nz[mw]) mrow = p.mask + ((size_t)mw * AN + i) * AN;
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    TMARK(0);
    uint32_t s[32], dp[32];
    tmem_ld32(tS + lane_off + wloc * 64 + jbase, s);
    tmem_ld32(tdP + lane_off + wloc * 64 + jbase, dp);
    tmem_ld_wait();
    TMARK(1);
    // P for this thread's 32 columns (fp32, registers) and the partial D = sum_j P_ij dP_ij, taken from the SAME
    // P and dP that form dS so that sum_j dS_ij == 0 up to fp32 rounding
    // (rows >= 49 / a missing window run on harmless finite values: their dO and Q rows are zero, so they add nothing to
    //  dV / dK, their dQ rows are never stored and their bias gradients are never written; columns >= 49 carry kNegBig)
    float pr[32];
    float delta = 0.f;
    {
      const float4* b4 = reinterpret_cast<const float4*>(sBias + (i < AN ? i : AN - 1) * kBiasLd + jbase);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 bb = b4[c];
        pr[4 * c + 0] = fmaf(__uint_as_float(s[4 * c + 0]), sc2, bb.x);
        pr[4 * c + 1] = fmaf(__uint_as_float(s[4 * c + 1]), sc2, bb.y);
        pr[4 * c + 2] = fmaf(__uint_as_float(s[4 * c + 2]), sc2, bb.z);
        pr[4 * c + 3] = fmaf(__uint_as_float(s[4 * c + 3]), sc2, bb.w);
      }
    }
    if (mrow != nullptr) {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        if (jbase + jj < AN) pr[jj] = fmaf(__ldg(mrow + jbase + jj), kLog2e, pr[jj]);
    }
    if (mb != 0u) {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        if ((mb >> jj) & 1u) pr[jj] -= 100.0f * kLog2e;
    }
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      const float pv = ex2_ftz(pr[jj] - lse2);
      delta = fmaf(pv, __uint_as_float(dp[jj]), delta);
      pr[jj] = pv;
    }
    sRed[hf][r] = delta;
    TMARK(2);
    if (warp == 0) {                             // previous item's dQ/dK/dV tiles (staged in sP/sdS) drained
      if (elect_one()) tma_store_wait_read<0>();
      __syncwarp();
    }
    __syncthreads();
    TMARK(3);
    delta = sRed[0][r] + sRed[1][r];
    float dsv[32];
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      const float ds = pr[jj] * (__uint_as_float(dp[jj]) - delta);
      db[jj] += ds;
      dsv[jj] = ds * p.scale;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      store_row_bf16x8(sP + sw128_off(r, hf * 4 + c), pr + 8 * c);
      store_row_bf16x8(sdS + sw128_off(r, hf * 4 + c), dsv + 8 * c);
    }
    TMARK(4);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    TMARK(5);
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk)   // dV[(w,j), (w',d)] = sum_i P_w[i][j] dO_w'[i][d]
          umma_bf16(tdV, umma_desc_join(kHi128, p_lo_mn + 128 * kk), umma_desc_join(kHi64, do_lo_mn + 64 * kk), idesc_tt, kk);
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk)   // dK = (scale dS)^T Q
          umma_bf16(tdK, umma_desc_join(kHi128, ds_lo_mn + 128 * kk), umma_desc_join(kHi64, q_lo_mn + 64 * kk), idesc_tt, kk);
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk)   // dQ = (scale dS) K
          umma_bf16(tdQ, umma_desc_join(kHi128, ds_lo_k + 2 * kk), umma_desc_join(kHi64, k_lo_mn + 64 * kk), idesc_nt, kk);
        umma_commit(bar_o);
      }
      __syncwarp();
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
    TMARK(6);
    {
      // dQ / dK / dV rows -> three bf16 tiles (8 KB each, 64-byte swizzle) in the sP + sdS region, which is free now
      // that the three MMAs have completed; then 49x32 TMA stores into the [q|k|v] column blocks of dqkv
      const uint32_t swz = (uint32_t)((r >> 1) & 3);
#pragma unroll
      for (int part = 0; part < 3; ++part) {   // 0: dQ, 1: dK, 2: dV
        uint32_t o[16];
        tmem_ld16((part == 0 ? tdQ : part == 1 ? tdK : tdV) + lane_off + wloc * 32 + hf * 16, o);
        tmem_ld_wait();
        uint8_t* trow = sP + part * 8192 + r * 64;
        store_row_bf16x8(trow + (((uint32_t)(hf * 2 + 0) ^ swz) << 4), reinterpret_cast<const float*>(o));
        store_row_bf16x8(trow + (((uint32_t)(hf * 2 + 1) ^ swz) << 4), reinterpret_cast<const float*>(o + 8));
      }
    }
    TMARK(7);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    TMARK(8);
    if (warp == 0) {
      if (elect_one()) {
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          if (2 * pair + w >= p.B_) continue;
#pragma unroll
          for (int part = 0; part < 3; ++part)
            tma_store_2d(&tmDQKV, aP + part * 8192 + w * 4096, part * p.C + h * AHD, (2 * pair + w) * AN);
        }
        tma_store_commit();
      }
      __syncwarp();
    }
  }
  TPRINT("attn_bwd");
  if (warp == 0) {
    if (elect_one()) tma_store_wait_all<0>();
    __syncwarp();
  }
  if (i < AN) {
    float* dst = p.dbias + ((size_t)h * AN + i) * AN + jbase;
#pragma unroll
    for (int jj = 0; jj < 32; ++jj)
      if (jbase + jj < AN) atomicAdd(dst + jj, db[jj]);
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
  int per_head = (persistent_sms() * ctas_per_sm) / a->nH;     // floor: the whole grid must be co-resident (one wave, no tail CTA)
  if (per_head > p.npairs) per_head = p.npairs;
  if (per_head < 1) per_head = 1;
  p.ctas_per_head = per_head;
  p.bias = a->bias; p.mask = a->mask; p.mask_nz = a->mask ? a->mask_nz : nullptr;
  p.canon_nwh = 0; p.canon_nww = 0;
  if (a->mask && a->canon_nwh > 0 && a->canon_nww > 0) {
    SWIN_REQUIRE(a->canon_nwh * a->canon_nww == a->nW, "attn: canonical mask grid %d x %d does not match nW = %d", a->canon_nwh, a->canon_nww, a->nW);
    p.canon_nwh = a->canon_nwh; p.canon_nww = a->canon_nww;
  }
  p.out = (__nv_bfloat16*)a->out; p.lse = a->lse;
  p.dout = (const __nv_bfloat16*)a->dout; p.dqkv = (__nv_bfloat16*)a->dqkv; p.dbias = a->dbias;
  *out = p;
  return 0;
}

int attn_tc_fwd(const swin_attn_args* a, cudaStream_t st) {
  AttnTcParams p;
  int rc = attn_tc_common(a, &p, false, 4);
  if (rc) return rc;
  if (p.B_ == 0) return 0;
  CUtensorMap tm;
  rc = make_tmap_bf16_2d(&tm, a->qkv, (uint64_t)3 * p.C, (uint64_t)p.B_ * AN, (uint64_t)3 * p.C * 2, AHD, AN, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  const size_t smem = kFwdTiles + kPBytes + (AN * kFwdBiasLd + 16) * sizeof(float) + 1024;
  rc = ensure_dyn_smem((const void*)attn_tc_fwd_kernel, (int)smem);
  if (rc) return rc;
  CUtensorMap tmo;
  rc = make_tmap_bf16_2d(&tmo, a->out, (uint64_t)p.C, (uint64_t)p.B_ * AN, (uint64_t)p.C * 2, AHD, AN, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  attn_tc_fwd_kernel<<<p.nH * p.ctas_per_head, kFwdThreads, smem, st>>>(tm, tmo, p);
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
  const size_t smem = 2 * kBwdTiles + 2 * kPBytes + AN * kBiasLd * sizeof(float) + 1024;
  rc = ensure_dyn_smem((const void*)attn_tc_bwd_kernel, (int)smem);
  if (rc) return rc;
  CUtensorMap tmdq;
  rc = make_tmap_bf16_2d(&tmdq, a->dqkv, (uint64_t)3 * p.C, (uint64_t)p.B_ * AN, (uint64_t)3 * p.C * 2, AHD, AN, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  attn_tc_bwd_kernel<<<p.nH * p.ctas_per_head, kAttnThreads, smem, st>>>(tm, tmdo, tmdq, p);
  SWIN_LAUNCH_CHECK();
  return 0;
}

}  // namespace swin
