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
#include <type_traits>
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


// Canonical SW-MSA mask (REF:370-389) in closed form, window 7 / shift 3.  Region ids on the padded grid differ only inside
// the last window row (R) / column (Cw) of the grid, where tokens with row (col) index >= ws - shift = 4 belong to another
// region than those < 4:   mask[i][j] = -100  <=>  (R and rowhi(j) != rowhi(i)) or (Cw and colhi(j) != colhi(i)).
// For an unrolled key column j, (rowhi(j), colhi(j)) is a compile-time CLASS (4 classes), so a row needs only four
// constants pen[rowhi][colhi] in {0, -100 log2(e)} -- and since every use of the masked logit is "logit + constant"
// (minus the row maximum / minus the saved LSE), the mask folds into that constant and costs no per-element instruction.
struct MaskPen { float c[2][2]; };
__device__ __forceinline__ MaskPen canon_mask_pen(bool active, int wi, int nwh, int nww, int i) {
  MaskPen m;
  const float kPen = -100.0f * 1.4426950408889634f;
  bool R = false, Cw = false;
  if (active) {
    const int wh = wi / nww, ww = wi - wh * nww;
    R = wh == nwh - 1; Cw = ww == nww - 1;
  }
  const int ii = i < AN ? i : AN - 1;
  const bool ri = ii / 7 >= 4, ci = ii % 7 >= 4;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) m.c[a][b] = ((R && ((a != 0) != ri)) || (Cw && ((b != 0) != ci))) ? kPen : 0.f;
  return m;
}
// class of key column j (compile-time in the unrolled loops); pad columns (j >= 49) fall in class (1, *), their bias is kNegBig
#define MASK_RH(j) (((j) / 7) >= 4 ? 1 : 0)
#define MASK_CH(j) (((j) % 7) >= 4 ? 1 : 0)

// Thread mapping of the backward kernel: 256 threads = 8 warps.  Warp w owns TMEM lane quarter q = w & 3 (rows
// q*32..q*32+31 of the stacked tile) and column half hf = w >> 2 of the row's own 64-column window block, so two
// threads cooperate on one row (row statistics are exchanged through smem).  This doubles the warps available to
// hide the LDS / TMEM / MUFU latencies of the per-row softmax math.
constexpr int kBiasLd = 68;                      // sBias row pitch (floats): 64 columns + 4 so that 8 consecutive rows' 16-byte
                                                 // reads fall in distinct bank groups; pre-scaled by log2(e); columns >= 49 hold
                                                 // kNegBig so padded keys drop out of the softmax with no per-element predicate
constexpr float kNegBig = -1.0e30f;
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// bias tile of head h -> shared memory, row pitch LD, pre-scaled by log2(e), pad columns kNegBig.  The global loads of a batch of
// four elements per thread are issued before any of them is used: as a plain load-then-store loop the prologue of a CTA was
// 13-20 dependent L2 round trips (~5 us of a 45-65 us launch at the stage-2/3 shapes).
template <int LD>
__device__ __forceinline__ void load_bias_tile_ld(float* sBias, const float* __restrict__ bias, int h) {
  const float kLog2e = 1.4426950408889634f;
  const float* src = bias + (size_t)h * AN * AN;
  const int n = AN * LD, step = blockDim.x;
  for (int e0 = threadIdx.x; e0 < n; e0 += 4 * step) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * step;
      const int i = e / LD, j = e - i * LD;
      v[u] = (e < n && j < AN) ? __ldg(src + i * AN + j) * kLog2e : kNegBig;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (e0 + u * step < n) sBias[e0 + u * step] = v[u];
  }
}
__device__ __forceinline__ void load_bias_tile(float* sBias, const float* __restrict__ bias, int h) { load_bias_tile_ld<kBiasLd>(sBias, bias, h); }

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
constexpr int kFwdPrefetch = 0;                    // items prefetched into L2 beyond the one whose tiles are being loaded
constexpr int kFwdBiasLd = 52;                     // bias row pitch of the forward kernel: 49 columns + 3 x kNegBig

__global__ void __launch_bounds__(kFwdThreads, 4) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                      const __grid_constant__ CUtensorMap tmOut, AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[5];              // qk, v, s, o, free (the staged O tile of the previous item has been read)
  __shared__ uint32_t tmem_slot;
  uint8_t* sbase = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sT = sbase;                                   // {Q,K,V}
  uint8_t* sP = sT + kFwdTiles;                          // 16 KB: P, later the O staging tile
  float* sBias = reinterpret_cast<float*>(sP + kPBytes); // [49][52]
  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x % p.nH, g = blockIdx.x / p.nH;
  const uint32_t bar_qk = smem_u32(&bars[0]), bar_v = smem_u32(&bars[1]), bar_s = smem_u32(&bars[2]), bar_o = smem_u32(&bars[3]), bar_free = smem_u32(&bars[4]);

  zero_smem(sT, kFwdTiles + kPBytes);
  if (tid == 0) {
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1); mbar_init(bar_free, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 128); tmem_relinquish(); }
  pdl_wait();                                    // (PDL) the above overlapped the previous kernel's tail; from here on global memory
  pdl_trigger();
  load_bias_tile_ld<kFwdBiasLd>(sBias, p.bias, h);
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
  // L2 prefetch of a later item's q, k, v boxes: the shared-memory tiles hold one item, so the HBM requests that keep the
  // memory system busy (bandwidth = bytes in flight / latency; tools/probes/tma_probe.cu) are issued ahead of the ring
  auto prefetch_item = [&](int pair) {
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int row0 = (2 * pair + w) * AN;
#pragma unroll
      for (int part = 0; part < 3; ++part) tma_prefetch_2d(&tmQKV, part * p.C + h * AHD, row0);
    }
  };
  // Single-thread work (TMA / MMA issue) sits on the item's serial chain: warp 0 enters converged and ONE elected lane issues
  // (no per-active-lane retry loops in the SASS), with descriptors built as constant-hi : incremented-lo words.
  constexpr uint32_t kHi64 = umma_desc_hi(512, (uint32_t)kSw64), kHi128 = umma_desc_hi(1024, (uint32_t)kSw128);
  const uint32_t q_lo = umma_desc_lo(aQ, 16), k_lo = umma_desc_lo(aK, 16), p_lo = umma_desc_lo(aP, 16), v_lo = umma_desc_lo(aV, 4096);
  auto issue_s = [&](uint32_t phase) {           // S = Q K^T as soon as Q, K have landed
    mbar_wait(bar_qk, phase);
    tc_fence_after();
#pragma unroll
    for (uint32_t k = 0; k < 2; ++k)
      umma_bf16(tS, umma_desc_join(kHi64, q_lo + 2 * k), umma_desc_join(kHi64, k_lo + 2 * k), idesc_s, k);
    umma_commit(bar_s);
  };
  // issue duties on two warps (as in the backward kernel): warp 0 every MMA, warp kTmaWarp every TMA load / store and the
  // bulk-group wait that belongs to the storing thread
  constexpr int kTmaWarp = 2;
  if (warp == kTmaWarp) {
    if (elect_one() && g < p.npairs) {
      issue_qk(g); issue_v(g);
      for (int d = 1; d <= kFwdPrefetch; ++d)
        if (g + d * stride < p.npairs) prefetch_item(g + d * stride);
    }
    __syncwarp();
  }

  uint32_t it = 0;
  for (int pair = g; pair < p.npairs; pair += stride, ++it) {
    const uint32_t ph = it & 1;
    const bool has_next = pair + stride < p.npairs;
    if (it == 0 && warp == 0) {                  // S = Q K^T of the first item (later items: issued at the end of the previous one)
      if (elect_one()) issue_s(0);
      __syncwarp();
    }
    const int win = 2 * pair + wloc;
    const bool valid = (i < AN) && (win < p.B_);
    const float* mrow = nullptr;
    bool canon = false;
    if (p.mask != nullptr && valid) {
      if (p.canon_nwh > 0) canon = true;
      else if (p.mask_nz == nullptr || p.mask_nz[win % p.nW]) mrow = p.mask + ((size_t)(win % p.nW) * AN + i) * AN;
    }
    const MaskPen pen = canon_mask_pen(canon, canon ? win % p.nW : 0, p.canon_nwh, p.canon_nww, i);
    mbar_wait(bar_s, ph);
    tc_fence_after();
    if (warp == kTmaWarp) {                      // the Q, K tiles are free again
      if (elect_one()) {
        if (has_next) issue_qk(pair + stride);
        if (pair + (kFwdPrefetch + 1) * stride < p.npairs) prefetch_item(pair + (kFwdPrefetch + 1) * stride);
      }
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
      const f32x2 sc = pk2(sc2, sc2);
#pragma unroll
      for (int c = 0; c < 13; ++c) {
        const float4 bb = b4[c];
        unpk2(fma2(pk2u(v[4 * c + 0], v[4 * c + 1]), sc, pk2(bb.x, bb.y)), sv[4 * c + 0], sv[4 * c + 1]);
        unpk2(fma2(pk2u(v[4 * c + 2], v[4 * c + 3]), sc, pk2(bb.z, bb.w)), sv[4 * c + 2], sv[4 * c + 3]);
      }
    }
    if (mrow != nullptr) {
#pragma unroll
      for (int jj = 0; jj < AN; ++jj) sv[jj] = fmaf(__ldg(mrow + jj), kLog2e, sv[jj]);
    }
    // row maximum of the MASKED logits from the four class maxima; exponent offset per class = penalty - maximum
    float cm[2][2] = {{sv[0], sv[4]}, {sv[28], sv[32]}};
#pragma unroll
    for (int jj = 1; jj < AN; ++jj) cm[MASK_RH(jj)][MASK_CH(jj)] = fmaxf(cm[MASK_RH(jj)][MASK_CH(jj)], sv[jj]);
    const float mx = fmaxf(fmaxf(cm[0][0] + pen.c[0][0], cm[0][1] + pen.c[0][1]), fmaxf(cm[1][0] + pen.c[1][0], cm[1][1] + pen.c[1][1]));
    const float off[2][2] = {{pen.c[0][0] - mx, pen.c[0][1] - mx}, {pen.c[1][0] - mx, pen.c[1][1] - mx}};
    float sum = 0.f, sum1 = 0.f;
#pragma unroll
    for (int jj = 0; jj < 52; jj += 2) {
      const float e0 = ex2_ftz(sv[jj] + off[MASK_RH(jj)][MASK_CH(jj)]);
      const float e1 = ex2_ftz(sv[jj + 1] + off[MASK_RH(jj + 1)][MASK_CH(jj + 1)]);
      sum += e0; sum1 += e1;
      sv[jj] = e0; sv[jj + 1] = e1;
    }
    sum += sum1;
    // the previous item's O tile is staged where P goes: its TMA store (issued a softmax ago) must have read it.  The wait
    // belongs to the thread that issued the store; everybody else learns through bar_free.
    if (warp == kTmaWarp) {
      if (elect_one()) { tma_store_wait_read<0>(); mbar_arrive(bar_free); }
      __syncwarp();
    }
    mbar_wait(bar_free, ph);
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
    if (warp == kTmaWarp) {                      // the V tile is free again
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
    if (warp == 0 && has_next) {
      // the next item's S MMA (every thread has read this item's O out of the TMEM columns it overwrites): the MMA round trip
      // is on the next item's chain
      if (elect_one()) issue_s(ph ^ 1);
      __syncwarp();
    }
    if (warp == kTmaWarp) {
      if (elect_one()) {
#pragma unroll
        for (int w = 0; w < 2; ++w)
          if (2 * pair + w < p.B_) tma_store_2d(&tmOut, aP + w * 4096, h * AHD, (2 * pair + w) * AN);
        tma_store_commit();
      }
      __syncwarp();
    }
  }
  if (warp == kTmaWarp) {
    if (elect_one()) tma_store_wait_all<0>();
    __syncwarp();
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_slot, 128); }
}

// ------------------------------------------------------------------------------------------ backward
// 256-thread CTAs (two threads per row of the 128-row tile), two per SM, double-buffered operand tiles.  What shapes the loop:
//   * dQ / dK / dV are staged IN PLACE over the item's own Q / K / V tiles (dead once its last MMAs have completed) and
//     TMA-stored from there, so the P / dS tiles are free for the next item at once.  The wait for those stores to have read
//     the buffer (cp.async.bulk.wait_group.read: ~1.5k cycles after the stores are issued) sits where it costs nothing --
//     one barrier into the NEXT item, right before that buffer is refilled with the item after next.
//   * dV = P^T dO is issued as soon as P is in shared memory and runs under the dS math.
//   * the saved LSE (HBM) is fetched one item ahead; the canonical mask costs four constants per row (canon_mask_pen).
//   (a TMA warp per CTA, or one 512/576-thread CTA per SM with two pipelines, were measured slower: registers are allocated
//    per CTA in 4-warp granules, so 9 or 18 warps drop the kernel to 96 registers and it spills.)
constexpr uint32_t kBwdTiles = 4 * kTileBytes;     // Q,K,V,dO per buffer
constexpr int kBwdThreads = 256;

__global__ void __launch_bounds__(kBwdThreads, 2) attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                      const __grid_constant__ CUtensorMap tmDO,
                                                                      const __grid_constant__ CUtensorMap tmDQKV, AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[5];               // load[2], s, o, v
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
  const uint32_t bar_load0 = smem_u32(&bars[0]), bar_s = smem_u32(&bars[2]), bar_o = smem_u32(&bars[3]), bar_v = smem_u32(&bars[4]);

  zero_smem(sT, 2 * kBwdTiles + 2 * kPBytes);
  if (tid == 0) {
    mbar_init(bar_load0, 1); mbar_init(bar_load0 + 8, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1); mbar_init(bar_v, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQKV);
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 256); tmem_relinquish(); }
  pdl_wait();                                    // (PDL) the above overlapped the previous kernel's tail; from here on global memory
  pdl_trigger();
  load_bias_tile(sBias, p.bias, h);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t aT = smem_u32(sT);
  const int stride = p.ctas_per_head;
  const uint32_t tS = tmem_slot, tdP = tmem_slot + 128;
  const uint32_t tdV = tmem_slot, tdK = tmem_slot + 64, tdQ = tmem_slot + 128;   // reuse S / dP columns after the softmax pass
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  const int r = q * 32 + lane, wloc = r >> 6, i = r & 63;
  const int jbase = hf * 32;
  const uint32_t idesc_s = umma_idesc_bf16(128, false, false);
  const uint32_t idesc_tt = umma_idesc_bf16(64, true, true);     // A^T B, both MN-major
  const uint32_t idesc_nt = umma_idesc_bf16(64, false, true);
  const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
  const float kLog2e = 1.4426950408889634f;
  const float sc2 = p.scale * kLog2e;
  constexpr uint32_t kHi64 = umma_desc_hi(512, (uint32_t)kSw64), kHi128 = umma_desc_hi(1024, (uint32_t)kSw128);
  const uint32_t ds_lo_k = umma_desc_lo(adS, 16);                                               // K-major view of dS
  const uint32_t p_lo_mn = umma_desc_lo(aP, 8192), ds_lo_mn = umma_desc_lo(adS, 8192);          // MN-major (transposed) views

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
  // Issue duties are split over two warps so that neither falls far behind the other seven between two block barriers: warp 0
  // issues every MMA, warp kTmaWarp every TMA load and store (bulk-group waits belong to the issuing thread, so loads and
  // stores stay together).
  constexpr int kTmaWarp = 4;
  if (warp == kTmaWarp) {
    if (elect_one()) {
      if (g < p.npairs) issue_loads(g, 0);
      if (g + stride < p.npairs) issue_loads(g + stride, 1);
    }
    __syncwarp();
  }

  // S = Q K^T and dP = dO V^T of the item in buffer `buf` (its `n`-th use)
  auto issue_sdp = [&](uint32_t buf, uint32_t n) {
    const uint32_t bQ = aT + buf * kBwdTiles;
    const uint32_t q_lo = umma_desc_lo(bQ, 16), k_lo = umma_desc_lo(bQ + kTileBytes, 16), v_lo = umma_desc_lo(bQ + 2 * kTileBytes, 16),
                   do_lo = umma_desc_lo(bQ + 3 * kTileBytes, 16);
    mbar_wait(bar_load0 + 8 * buf, n & 1);
    tc_fence_after();
#pragma unroll
    for (uint32_t k = 0; k < 2; ++k)
      umma_bf16(tS, umma_desc_join(kHi64, q_lo + 2 * k), umma_desc_join(kHi64, k_lo + 2 * k), idesc_s, k);
#pragma unroll
    for (uint32_t k = 0; k < 2; ++k)
      umma_bf16(tdP, umma_desc_join(kHi64, do_lo + 2 * k), umma_desc_join(kHi64, v_lo + 2 * k), idesc_s, k);
    umma_commit(bar_s);
  };

  f32x2 db2[16];                                   // dBias partial sums of this thread's (row, 32 columns), packed pairs
#pragma unroll
  for (int j = 0; j < 16; ++j) db2[j] = pk2(0.f, 0.f);
  TDECL
  // the saved LSE of a row comes from HBM: fetched one item ahead so that its latency never sits on the item's chain
  auto load_lse = [&](int pair) -> float {
    const int win = 2 * pair + wloc;
    return ((i < AN) && (win < p.B_)) ? p.lse[((size_t)win * p.nH + h) * AN + i] : 0.f;      // raw: see the use below
  };
  float lse_next = g < p.npairs ? load_lse(g) : 0.f;
  const f32x2 sc2p = pk2(sc2, sc2), scalep = pk2(p.scale, p.scale);
  const float4* brow = reinterpret_cast<const float4*>(sBias + (i < AN ? i : AN - 1) * kBiasLd + jbase);

  uint32_t it = 0;
  for (int pair = g; pair < p.npairs; pair += stride, ++it) {
    const uint32_t ph = it & 1, buf = it & 1;
    const uint32_t aQ = aT + buf * kBwdTiles, aK = aQ + kTileBytes, adO = aQ + 3 * kTileBytes;
    TITEM
    // MN-major views of this buffer's Q, K, dO tiles as descriptor low words (LBO 4096)
    const uint32_t q_lo_mn = umma_desc_lo(aQ, 4096), k_lo_mn = umma_desc_lo(aK, 4096), do_lo_mn = umma_desc_lo(adO, 4096);
    const bool has_next = pair + stride < p.npairs;
    if (it == 0 && warp == 0) {                    // (later items: issued at the end of the previous item, ahead of its stores)
      if (elect_one()) issue_sdp(0, 0);
      __syncwarp();
    }
    const int win = 2 * pair + wloc;
    const bool valid = (i < AN) && (win < p.B_);
    // the value fetched during the PREVIOUS item is scaled only here: a multiply placed next to the load made every warp wait
    // for the HBM round trip on the spot (13.6 % of the kernel's stall samples, profiles/r02/r02_attn_bwd_s0.hot.txt)
    const float lse2 = lse_next * kLog2e;
    if (has_next) lse_next = load_lse(pair + stride);
    const float* mrow = nullptr;
    bool canon = false;
    if (p.mask != nullptr && valid) {
      if (p.canon_nwh > 0) canon = true;
      else if (p.mask_nz == nullptr || p.mask_nz[win % p.nW]) mrow = p.mask + ((size_t)(win % p.nW) * AN + i) * AN;
    }
    // exponent offset per mask class: penalty - LSE (see canon_mask_pen)
    MaskPen off = canon_mask_pen(canon, canon ? win % p.nW : 0, p.canon_nwh, p.canon_nww, i);
    off.c[0][0] -= lse2; off.c[0][1] -= lse2; off.c[1][0] -= lse2; off.c[1][1] -= lse2;
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
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 bb = brow[c];
      unpk2(fma2(pk2u(s[4 * c + 0], s[4 * c + 1]), sc2p, pk2(bb.x, bb.y)), pr[4 * c + 0], pr[4 * c + 1]);
      unpk2(fma2(pk2u(s[4 * c + 2], s[4 * c + 3]), sc2p, pk2(bb.z, bb.w)), pr[4 * c + 2], pr[4 * c + 3]);
    }
    if (mrow != nullptr) {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        if (jbase + jj < AN) pr[jj] = fmaf(__ldg(mrow + jbase + jj), kLog2e, pr[jj]);
    }
    float delta = 0.f, delta1 = 0.f;
    // the mask class of column jbase + jj is a compile-time constant per column half
    auto exp_half = [&](auto HF) {
      constexpr int jb = decltype(HF)::value * 32;
#pragma unroll
      for (int jj = 0; jj < 32; jj += 2) {
        const float p0 = ex2_ftz(pr[jj] + off.c[MASK_RH(jb + jj)][MASK_CH(jb + jj)]);
        const float p1 = ex2_ftz(pr[jj + 1] + off.c[MASK_RH(jb + jj + 1)][MASK_CH(jb + jj + 1)]);
        delta = fmaf(p0, __uint_as_float(dp[jj]), delta);
        delta1 = fmaf(p1, __uint_as_float(dp[jj + 1]), delta1);
        pr[jj] = p0; pr[jj + 1] = p1;
      }
    };
    if (hf == 0) exp_half(std::integral_constant<int, 0>{}); else exp_half(std::integral_constant<int, 1>{});
    delta += delta1;
    sRed[hf][r] = delta;
#pragma unroll
    for (int c = 0; c < 4; ++c) store_row_bf16x8(sP + sw128_off(r, hf * 4 + c), pr + 8 * c);
    TMARK(2);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    TMARK(3);
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk)   // dV[(w,j), (w',d)] = sum_i P_w[i][j] dO_w'[i][d]  -- runs under the dS math
          umma_bf16(tdV, umma_desc_join(kHi128, p_lo_mn + 128 * kk), umma_desc_join(kHi64, do_lo_mn + 64 * kk), idesc_tt, kk);
        umma_commit(bar_v);                     // dV has its own barrier: it is staged while the dK / dQ MMAs are still running
      }
      __syncwarp();
    }
    if (warp == kTmaWarp && it >= 1 && has_next) {
      // the PREVIOUS item's dQ / dK / dV (staged over its own tiles, stores issued ~2k cycles ago) have been read out: refill
      // that buffer with the next item
      if (elect_one()) { tma_store_wait_read<0>(); issue_loads(pair + stride, buf ^ 1); }
      __syncwarp();
    }
    delta = sRed[0][r] + sRed[1][r];
    {
      const f32x2 nd = pk2(-delta, -delta);
      float dsv[32];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const f32x2 ds = mul2(pk2(pr[2 * c], pr[2 * c + 1]), add2(pk2u(dp[2 * c], dp[2 * c + 1]), nd));
        db2[c] = add2(db2[c], ds);
        unpk2(mul2(ds, scalep), dsv[2 * c], dsv[2 * c + 1]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) store_row_bf16x8(sdS + sw128_off(r, hf * 4 + c), dsv + 8 * c);
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
        for (uint32_t kk = 0; kk < 4; ++kk)   // dK = (scale dS)^T Q
          umma_bf16(tdK, umma_desc_join(kHi128, ds_lo_mn + 128 * kk), umma_desc_join(kHi64, q_lo_mn + 64 * kk), idesc_tt, kk);
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk)   // dQ = (scale dS) K
          umma_bf16(tdQ, umma_desc_join(kHi128, ds_lo_k + 2 * kk), umma_desc_join(kHi64, k_lo_mn + 64 * kk), idesc_nt, kk);
        umma_commit(bar_o);
      }
      __syncwarp();
    }
    {
      // dQ / dK / dV rows -> bf16, staged IN PLACE over this item's Q / K / V tiles (64-byte swizzle, window w at +4096) once every
      // MMA that read the tile has completed.  Pad rows (i >= 49) are written as zeros so the tiles stay valid operand tiles when
      // the refill rewrites rows 0..48 only.  dV goes first: its MMAs finished under the dS math and the V tile was last read by
      // the dP MMA, so this third of the epilogue runs while the dK / dQ MMAs (which still read Q and K) are in flight.
      const uint32_t swz = (uint32_t)((r >> 1) & 3);
      uint8_t* tbase = sT + buf * kBwdTiles + r * 64;
      auto put = [&](uint32_t* o, int part) {
        if (i >= AN) {
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = 0u;
        }
        uint8_t* trow = tbase + part * kTileBytes;
        store_row_bf16x8(trow + (((uint32_t)(hf * 2 + 0) ^ swz) << 4), reinterpret_cast<const float*>(o));
        store_row_bf16x8(trow + (((uint32_t)(hf * 2 + 1) ^ swz) << 4), reinterpret_cast<const float*>(o + 8));
      };
      uint32_t o[32];
      mbar_wait(bar_v, ph);
      tc_fence_after();
      tmem_ld16(tdV + lane_off + wloc * 32 + hf * 16, o);
      tmem_ld_wait();
      put(o, 2);
      mbar_wait(bar_o, ph);
      tc_fence_after();
      TMARK(6);
      tmem_ld16(tdQ + lane_off + wloc * 32 + hf * 16, o);
      tmem_ld16(tdK + lane_off + wloc * 32 + hf * 16, o + 16);
      tmem_ld_wait();
      put(o, 0);
      put(o + 16, 1);
    }
    TMARK(7);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();                               // also: every warp has read dQ / dK / dV out of TMEM before the next item's S / dP MMAs
    TMARK(8);
    if (warp == 0 && has_next) {
      // the next item's S / dP MMAs (its tiles landed long ago; every warp has read this item's outputs out of the TMEM columns
      // they overwrite): the MMA round trip is the longest wait of an item
      if (elect_one()) issue_sdp(buf ^ 1, (it + 1) >> 1);
      __syncwarp();
    }
    if (warp == kTmaWarp) {
      if (elect_one()) {
        const uint32_t base = aT + buf * kBwdTiles;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          if (2 * pair + w >= p.B_) continue;
#pragma unroll
          for (int part = 0; part < 3; ++part)
            tma_store_2d(&tmDQKV, base + part * kTileBytes + w * 4096, part * p.C + h * AHD, (2 * pair + w) * AN);
        }
        tma_store_commit();
      }
      __syncwarp();
    }
  }
  TPRINT("attn_bwd");
  if (warp == kTmaWarp) {
    if (elect_one()) tma_store_wait_all<0>();
    __syncwarp();
  }
  // dBias: the two windows of the tile hold the same (i, j) entries -- rows r and r + 64 are summed through shared memory
  // (the P / dS tiles are dead: the last MMAs have completed) before the global atomics
  if (g < p.npairs) {
    float* mine = reinterpret_cast<float*>(sP) + ((hf * 64 + i) * 32);            // [2 column halves][64 rows][32] floats = 16 KB
    if (wloc == 1) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 t;
        unpk2(db2[2 * c], t.x, t.y); unpk2(db2[2 * c + 1], t.z, t.w);
        reinterpret_cast<float4*>(mine)[c ^ (i & 7)] = t;
      }
    }
    __syncthreads();
    if (wloc == 0 && i < AN) {
      float* dst = p.dbias + ((size_t)h * AN + i) * AN + jbase;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 t = reinterpret_cast<const float4*>(mine)[c ^ (i & 7)];
        float a0, a1, a2, a3;
        unpk2(db2[2 * c], a0, a1); unpk2(db2[2 * c + 1], a2, a3);
        if (jbase + 4 * c + 0 < AN) atomicAdd(dst + 4 * c + 0, a0 + t.x);
        if (jbase + 4 * c + 1 < AN) atomicAdd(dst + 4 * c + 1, a1 + t.y);
        if (jbase + 4 * c + 2 < AN) atomicAdd(dst + 4 * c + 2, a2 + t.z);
        if (jbase + 4 * c + 3 < AN) atomicAdd(dst + 4 * c + 3, a3 + t.w);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
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
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.nH * p.ctas_per_head)); cfg.blockDim = dim3(kFwdThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute lattr[1];
    cfg.attrs = lattr; cfg.numAttrs = pdl_attr(&lattr[0]);
    const cudaError_t le = cudaLaunchKernelEx(&cfg, attn_tc_fwd_kernel, tm, tmo, p);
    if (le != cudaSuccess) { set_error("attn_tc_fwd launch: %s", cudaGetErrorString(le)); return (int)le; }
  }
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
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.nH * p.ctas_per_head)); cfg.blockDim = dim3(kBwdThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute lattr[1];
    cfg.attrs = lattr; cfg.numAttrs = pdl_attr(&lattr[0]);
    const cudaError_t le = cudaLaunchKernelEx(&cfg, attn_tc_bwd_kernel, tm, tmdo, tmdq, p);
    if (le != cudaSuccess) { set_error("attn_tc_bwd launch: %s", cudaGetErrorString(le)); return (int)le; }
  }
  SWIN_LAUNCH_CHECK();
  return 0;
}

}  // namespace swin
