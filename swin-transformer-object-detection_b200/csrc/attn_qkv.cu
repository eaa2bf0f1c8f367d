// Fused QKV projection + window attention, forward (window 7, head_dim 32, bf16 operands, fp32 accumulation in TMEM).
//
//   REF = mmdet/models/backbones/swin_transformer.py:128-150:  qkv = x W^T + b ; S = scale q k^T + bias + mask ; P = softmax(S) ; O = P v
//
// One kernel replaces the qkv GEMM + the attention kernel: the LayerNorm'd window rows are read ONCE, Q / K / V of a head are
// produced by a tcgen05 GEMM against the (shared-memory resident) weight slice of that head, converted to bf16 operand tiles in
// shared memory and consumed by the S / softmax / P.V pipeline without ever touching HBM.  HBM traffic per window pair is
// 2 x 49 x C bf16 in + the same out (the stand-alone pair of kernels moves 4x that), which lifts the branch from
// 24.5 flop/B (attention alone) to ~190 flop/B at C = 96.
//
// Work item n = (window-pair tile k, head h), n = k * nH + h, walked in order by one CTA per SM.  Roles (384 threads):
//   warp 0      TMA producer: the weight slices once, then the X tile of every window pair into a 2-slot ring
//   warp 1 / 2  MMA issuer of group A / B (one elected lane each): per item  ACC = X W_h^T (128 x 96 x C)  ->
//               S = Q K^T (128 x 128 x 32)  ->  O = P [V0|V1] (128 x 64 x 64).  One issuer per group: a tcgen05.commit blocks its
//               thread for ~300 cycles and an item needs three, so a single issuer for both groups would be the bottleneck.
//   warp 3      store warp: TMA stores of the O tiles of both groups and the wait for the stores to have read their staging
//               tiles -- kept off the groups' critical path
//   warps 4-7   softmax group A: items n even        warps 8-11  softmax group B: items n odd
//               per item: ACC (+bias) -> bf16 Q, K, V tiles in smem | S -> P (registers, thread = row) | O -> bf16 staging tile
// The two groups alternate items, so while one group converts / stages, the other runs its softmax and the tensor core works
// for both (the item itself is a serial chain; a second CTA per SM does not fit next to the resident weights).  Hand-overs are
// mbarriers with one arrival per warp (no CTA- or group-wide bar.sync inside the loop).
// TMEM (512 columns): group g owns [256 g, 256 g + 256): ACC in [0, 96), S in [128, 256), O overlays S[0, 64).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace swin {
namespace {

constexpr int QHD = 32;            // head dim
constexpr int QN = 49;             // tokens per window (ws = 7)
constexpr int kQThreads = 384;
constexpr int kRelLd = 52;         // rel-bias row pitch (floats): 49 columns + 3 x kNegBigQ
constexpr float kNegBigQ = -1.0e30f;
constexpr size_t kMaxDynSmem = 227 * 1024 - 4096;     // opt-in limit minus head-room for the kernel's static shared memory
constexpr uint32_t kQkvTileBytes = 3 * 8192;    // Q, K, V operand tiles of one group: 128 rows x 32 bf16 each (SW64), window 1 at +4096
constexpr uint32_t kPTileBytes = 16384;         // compact P: 128 rows x 64 bf16 (SW128); later the O staging tile
constexpr uint32_t kXFull = 16384, kXTail = 8192;      // X k-blocks: 128 rows x 64 (SW128) / x 32 (SW64) bf16
constexpr uint32_t kWFull = 12288, kWTail = 6144;      // W_h k-blocks: 96 rows x 64 (SW128) / x 32 (SW64) bf16

struct AttnQkvParams {
  int B_, nH, nW, C, ntiles;
  int nfull, tail;                 // C = 64 * nfull + 32 * tail
  int canon_nwh, canon_nww;
  float scale;
  const float* rel_bias; const float* mask; const int* mask_nz; const float* bqkv;
  float* lse;
  __nv_bfloat16* qkv_out;
  uint32_t x_slot_bytes, w_head_bytes;
};

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t sw128o(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ void st_bf16x8(uint8_t* dst, const float* v) {
  int4 pk;
  pk.x = pack_bf16(v[0], v[1]); pk.y = pack_bf16(v[2], v[3]); pk.z = pack_bf16(v[4], v[5]); pk.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<int4*>(dst) = pk;
}

// canonical SW-MSA mask of one row in closed form (see attn_tc.cu): bit j = 1 <=> mask[w][i][j] == -100
__device__ __forceinline__ unsigned long long canon_bits(int wi, int nwh, int nww, int i) {
  constexpr unsigned long long kRowHi = 0x1FFFFF0000000ULL, kColHi = 0x1C3870E1C3870ULL, kAll = 0x1FFFFFFFFFFFFULL;
  const int wh = wi / nww, ww = wi - wh * nww;
  unsigned long long m = 0ULL;
  if (wh == nwh - 1) m |= (i / 7 >= 4) ? (~kRowHi & kAll) : kRowHi;
  if (ww == nww - 1) m |= (i % 7 >= 4) ? (~kColHi & kAll) : kColHi;
  return m;
}

// Development aid (-DSWIN_QKV_TIMING, never in the shipped library): one softmax thread of CTA 0 accumulates clock64() deltas per phase.
#ifdef SWIN_QKV_TIMING
#define QT_DECL long long qt[12] = {0}; long long qprev = clock64(); int qitems = 0;
#define QT(k) do { if (blockIdx.x == 0 && tid == 128 + 8) { long long t_ = clock64(); qt[k] += t_ - qprev; qprev = t_; } } while (0)
#define QT_ITEM ++qitems;
#define QT_PRINT do { if (blockIdx.x == 0 && tid == 128 + 8 && qitems) printf("attn_qkv items=%d clk/item: acc_wait %lld | ld+convert %lld | arrive qk %lld | mask prep %lld | s_wait %lld | ldS %lld | softmax %lld | Pst+arrive %lld | o_wait %lld | O epi %lld | arrive ost %lld\n", qitems, qt[0]/qitems, qt[1]/qitems, qt[2]/qitems, qt[3]/qitems, qt[4]/qitems, qt[5]/qitems, qt[6]/qitems, qt[7]/qitems, qt[8]/qitems, qt[9]/qitems, qt[10]/qitems); } while (0)
#else
#define QT_DECL
#define QT(k)
#define QT_ITEM
#define QT_PRINT
#endif

// barrier slots
enum { B_WFULL = 0, B_XFULL = 1, B_XEMPTY = 3, B_ACC = 5, B_QK = 7, B_S = 9, B_P = 11, B_O = 13, B_OST = 15, B_OFREE = 17, B_COUNT = 19 };

// wait of the utility warps (producer, store): same bounded wait, but backing off between polls so the spin does not take
// issue slots from the softmax warps on the same scheduler
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if ((++spins & 1023) == 0 && clock64() - t0 > 4000000000LL) { atomicExch(&g_watchdog_flag, 1); __trap(); }
  }
}

__global__ void __launch_bounds__(kQThreads, 1)
attn_qkv_fwd_kernel(const __grid_constant__ CUtensorMap tmX128, const __grid_constant__ CUtensorMap tmX64,
                    const __grid_constant__ CUtensorMap tmW128, const __grid_constant__ CUtensorMap tmW64,
                    const __grid_constant__ CUtensorMap tmOut, AttnQkvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[B_COUNT + 1];
  __shared__ uint32_t tmem_slot;
  uint8_t* sbase = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = sbase;                                              // nH x w_head_bytes
  uint8_t* sX = sW + (uint32_t)p.nH * p.w_head_bytes;               // 2 x x_slot_bytes
  uint8_t* sG = sX + 2u * p.x_slot_bytes;                           // 2 groups x {Q,K,V tiles, P tile}
  float* sRel = reinterpret_cast<float*>(sG + 2u * (kQkvTileBytes + kPTileBytes));   // nH x [49][52], pre-scaled by log2(e)
  float* sBq = sRel + p.nH * QN * kRelLd;                           // 3C qkv bias (zeros if the Linear has none)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const float kLog2e = 1.4426950408889634f;

  // ---- prologue: zero the X ring (its pad rows 49..63 are never written by TMA), stage rel-bias tiles and the qkv bias
  for (uint32_t o = tid * 16; o < 2u * p.x_slot_bytes; o += kQThreads * 16) *reinterpret_cast<int4*>(sX + o) = make_int4(0, 0, 0, 0);
  for (int e = tid; e < p.nH * QN * kRelLd; e += kQThreads) {
    const int hh = e / (QN * kRelLd), rem = e - hh * (QN * kRelLd), bi = rem / kRelLd, bj = rem - bi * kRelLd;
    sRel[e] = bj < QN ? p.rel_bias[((size_t)hh * QN + bi) * QN + bj] * kLog2e : kNegBigQ;
  }
  for (int e = tid; e < 3 * p.C; e += kQThreads) sBq[e] = p.bqkv != nullptr ? p.bqkv[e] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i) {
      const bool per_warp = (i >= B_QK && i < B_QK + 2) || (i >= B_P && i < B_P + 2) || (i >= B_OST && i < B_OST + 2);   // one arrival per group warp
      const bool per_issuer = (i >= B_XEMPTY && i < B_XEMPTY + 2) && p.nH >= 2;      // both groups' issuers use every tile
      mbar_init(bar(i), per_warp ? 4 : (per_issuer ? 2 : 1));
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmX128); tma_prefetch_desc(&tmW128); tma_prefetch_desc(&tmOut);
  }
  if (warp == 1) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  const int G = gridDim.x;
  const int nk = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - (int)blockIdx.x + G - 1) / G : 0;     // tiles of this CTA
  const int nitems = nk * p.nH;
  const uint32_t aW = smem_u32(sW), aX = smem_u32(sX), aG = smem_u32(sG);

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      mbar_expect_tx(bar(B_WFULL), (uint32_t)(3 * p.C * p.C * 2));
      for (int h = 0; h < p.nH; ++h) {
        const uint32_t wb = aW + (uint32_t)h * p.w_head_bytes;
        for (int kb = 0; kb < p.nfull; ++kb)
          for (int part = 0; part < 3; ++part) tma_load_2d(wb + kb * kWFull + part * 4096, &tmW128, bar(B_WFULL), kb * 64, part * p.C + h * QHD);
        if (p.tail)
          for (int part = 0; part < 3; ++part) tma_load_2d(wb + p.nfull * kWFull + part * 2048, &tmW64, bar(B_WFULL), p.nfull * 64, part * p.C + h * QHD);
      }
      for (int k = 0; k < nk; ++k) {
        const int slot = k & 1;
        if (k >= 2) mbar_wait_relaxed(bar(B_XEMPTY + slot), ((k >> 1) - 1) & 1);
        const uint32_t xb = aX + (uint32_t)slot * p.x_slot_bytes, fb = bar(B_XFULL + slot);
        mbar_expect_tx(fb, (uint32_t)(2 * QN * p.C * 2));
        const int tile = blockIdx.x + k * G;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int row0 = (2 * tile + w) * QN;             // rows past the tensor (odd window count) arrive as zeros
          for (int kb = 0; kb < p.nfull; ++kb) tma_load_2d(xb + kb * kXFull + w * 8192, &tmX128, fb, kb * 64, row0);
          if (p.tail) tma_load_2d(xb + p.nfull * kXFull + w * 4096, &tmX64, fb, p.nfull * 64, row0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2) {
    // ===================================================== MMA issuer of group g (one thread walks the group's items)
    const int g = warp - 1;
    if (elect_one()) {
      constexpr uint32_t kHi64 = umma_desc_hi(512, (uint32_t)kSw64), kHi128 = umma_desc_hi(1024, (uint32_t)kSw128);
      const uint32_t idesc_qkv = umma_idesc_bf16(96, false, false);
      const uint32_t idesc_s = umma_idesc_bf16(128, false, false);
      const uint32_t idesc_o = umma_idesc_bf16(64, false, true);
      const uint32_t tAcc = tmem + g * 256, tS = tAcc + 128;
      const uint32_t aQ = aG + g * (kQkvTileBytes + kPTileBytes), aK = aQ + 8192, aV = aQ + 16384, aP = aQ + kQkvTileBytes;
      const uint32_t qlo = umma_desc_lo(aQ, 16), klo = umma_desc_lo(aK, 16), plo = umma_desc_lo(aP, 16), vlo = umma_desc_lo(aV, 4096);
      int xk = -1;                                   // newest tile whose X slot this thread has seen filled
      // projection of item n; false (nothing issued) if its X tile has not arrived and `block` is false
      auto issue_qkv = [&](int n, bool block) -> bool {
        const int k = n / p.nH, h = n - k * p.nH, slot = k & 1;
        if (k > xk) {
          const uint32_t fb = bar(B_XFULL + slot), par = (uint32_t)(k >> 1) & 1;
          if (block) mbar_wait(fb, par); else if (!mbar_test_wait(fb, par)) return false;
          tc_fence_after();
          xk = k;
        }
        const uint32_t xb = aX + (uint32_t)slot * p.x_slot_bytes, wb = aW + (uint32_t)h * p.w_head_bytes;
        uint32_t acc = 0;
        for (int kb = 0; kb < p.nfull; ++kb) {
          const uint32_t xlo = umma_desc_lo(xb + kb * kXFull, 16), wlo = umma_desc_lo(wb + kb * kWFull, 16);
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks) { umma_bf16(tAcc, umma_desc_join(kHi128, xlo + 2 * ks), umma_desc_join(kHi128, wlo + 2 * ks), idesc_qkv, acc); acc = 1; }
        }
        if (p.tail) {
          const uint32_t xlo = umma_desc_lo(xb + p.nfull * kXFull, 16), wlo = umma_desc_lo(wb + p.nfull * kWFull, 16);
#pragma unroll
          for (uint32_t ks = 0; ks < 2; ++ks) { umma_bf16(tAcc, umma_desc_join(kHi64, xlo + 2 * ks), umma_desc_join(kHi64, wlo + 2 * ks), idesc_qkv, acc); acc = 1; }
        }
        umma_commit(bar(B_ACC + g));
        // the X slot is free once every head's projection has run: each issuer reports after ITS last item of the tile
        if (n + 2 >= (k + 1) * p.nH) umma_commit(bar(B_XEMPTY + slot));
        return true;
      };
      if (g < nitems) {
        mbar_wait(bar(B_WFULL), 0);
        tc_fence_after();
        issue_qkv(g, true);
      }
      for (int n = g; n < nitems; n += 2) {
        const uint32_t ph = (uint32_t)(n >> 1) & 1;
        mbar_wait(bar(B_QK + g), ph);                // Q, K, V tiles of item n are in smem and ACC has been read out
        tc_fence_after();
#pragma unroll
        for (uint32_t ks = 0; ks < 2; ++ks) umma_bf16(tS, umma_desc_join(kHi64, qlo + 2 * ks), umma_desc_join(kHi64, klo + 2 * ks), idesc_s, ks);
        umma_commit(bar(B_S + g));
        // the next item's projection runs under this item's softmax (if its X tile is already here; else after P.V)
        const bool more = n + 2 < nitems;
        const bool early = more && issue_qkv(n + 2, false);
        mbar_wait(bar(B_P + g), ph);
        tc_fence_after();
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk) umma_bf16(tS, umma_desc_join(kHi128, plo + 2 * kk), umma_desc_join(kHi64, vlo + 64 * kk), idesc_o, kk);
        umma_commit(bar(B_O + g));
        if (more && !early) issue_qkv(n + 2, true);
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================================================== store warp: O tiles of both groups, in item order
    if (elect_one()) {
      for (int n = 0; n < nitems; ++n) {
        const int g = n & 1;
        const uint32_t ph = (uint32_t)(n >> 1) & 1;
        const int k = n / p.nH, h = n - k * p.nH;
        const int tile = blockIdx.x + k * G;
        const uint32_t aP = aG + g * (kQkvTileBytes + kPTileBytes) + kQkvTileBytes;
        mbar_wait_relaxed(bar(B_OST + g), ph);
#pragma unroll
        for (int w = 0; w < 2; ++w)
          if (2 * tile + w < p.B_) tma_store_2d(&tmOut, aP + w * 4096, h * QHD, (2 * tile + w) * QN);
        tma_store_commit();
        if (n >= 1) {                                // one item of lag: the PREVIOUS item's stores have read their staging tile
          tma_store_wait_read<1>();
          mbar_arrive(bar(B_OFREE + (g ^ 1)));
        }
      }
      if (nitems > 0) { tma_store_wait_read<0>(); mbar_arrive(bar(B_OFREE + ((nitems - 1) & 1))); }
      tma_store_wait_all<0>();
    }
    __syncwarp();
  } else {
    // ===================================================== softmax groups (thread = one row of the stacked 128-row tile)
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int r = q * 32 + lane, wloc = r >> 6, i = r & 63;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tAcc = tmem + g * 256 + lane_off, tS = tAcc + 128;
    uint8_t* sQ = sG + g * (kQkvTileBytes + kPTileBytes);
    uint8_t* sP = sQ + kQkvTileBytes;
    const float sc2 = p.scale * kLog2e;
    const uint32_t swz = (uint32_t)((r >> 1) & 3);
    // hand-over to the MMA / store warps: every thread has fenced its smem writes (generic -> async proxy) and its TMEM reads,
    // then one lane per warp arrives (the barriers expect 4 arrivals)
    auto warp_arrive = [&](int b) {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(b));
    };

    QT_DECL
    int mask_k = -1;
    const float* mrow = nullptr;
    unsigned long long mb = 0ULL;
    for (int n = g; n < nitems; n += 2) {
      const uint32_t ph = (uint32_t)(n >> 1) & 1;
      QT_ITEM
      const int k = n / p.nH, h = n - k * p.nH;
      const int tile = blockIdx.x + k * G;
      const int win = 2 * tile + wloc;
      const bool valid = (i < QN) && (win < p.B_);
      // ---- (a) projection accumulator (+ bias) -> bf16 Q, K, V operand tiles (and, for training, q, k, v rows to HBM)
      mbar_wait(bar(B_ACC + g), ph);
      tc_fence_after();
      QT(0);
      {
        uint32_t v[96];
        tmem_ld32(tAcc, v);
        tmem_ld32(tAcc + 32, v + 32);
        tmem_ld32(tAcc + 64, v + 64);
        tmem_ld_wait();
        __nv_bfloat16* grow = (p.qkv_out != nullptr && valid) ? p.qkv_out + ((size_t)win * QN + i) * (3 * p.C) + h * QHD : nullptr;
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          const float4* b4 = reinterpret_cast<const float4*>(sBq + part * p.C + h * QHD);
          uint8_t* trow = sQ + part * 8192 + r * 64;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 b0 = b4[2 * c], b1 = b4[2 * c + 1];
            const uint32_t* vv = v + part * 32 + 8 * c;
            int4 pk;
            pk.x = pack_bf16(__uint_as_float(vv[0]) + b0.x, __uint_as_float(vv[1]) + b0.y);
            pk.y = pack_bf16(__uint_as_float(vv[2]) + b0.z, __uint_as_float(vv[3]) + b0.w);
            pk.z = pack_bf16(__uint_as_float(vv[4]) + b1.x, __uint_as_float(vv[5]) + b1.y);
            pk.w = pack_bf16(__uint_as_float(vv[6]) + b1.z, __uint_as_float(vv[7]) + b1.w);
            *reinterpret_cast<int4*>(trow + (((uint32_t)c ^ swz) << 4)) = pk;
            // each row's 64 bytes of q / k / v are two whole 32-byte sectors: written straight from the registers
            if (grow != nullptr) *reinterpret_cast<int4*>(grow + part * p.C + 8 * c) = pk;
          }
        }
      }
      QT(1);
      warp_arrive(B_QK + g);
      QT(2);
      // ---- (b) S -> P
      if (k != mask_k) {                              // the row's mask depends on the window only: once per tile, not per head
        mask_k = k; mrow = nullptr; mb = 0ULL;
        if (p.mask != nullptr && valid) {
          const int mw = win % p.nW;
          if (p.canon_nwh > 0) mb = canon_bits(mw, p.canon_nwh, p.canon_nww, i);
          else if (p.mask_nz == nullptr || p.mask_nz[mw]) mrow = p.mask + ((size_t)mw * QN + i) * QN;
        }
      }
      QT(3);
      mbar_wait(bar(B_S + g), ph);
      tc_fence_after();
      QT(4);
      uint32_t v[52];
      tmem_ld32(tS + wloc * 64, v);
      tmem_ld16(tS + wloc * 64 + 32, v + 32);
      tmem_ld4(tS + wloc * 64 + 48, v + 48);
      tmem_ld_wait();
      QT(5);
      float sv[52];
      {
        const float4* b4 = reinterpret_cast<const float4*>(sRel + (h * QN + (i < QN ? i : QN - 1)) * kRelLd);
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
        for (int jj = 0; jj < QN; ++jj) sv[jj] = fmaf(__ldg(mrow + jj), kLog2e, sv[jj]);
      }
      if (mb != 0ULL) {
        const uint32_t lo = (uint32_t)mb, hi = (uint32_t)(mb >> 32);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          if ((lo >> jj) & 1u) sv[jj] -= 100.0f * kLog2e;
#pragma unroll
        for (int jj = 32; jj < QN; ++jj)
          if ((hi >> (jj - 32)) & 1u) sv[jj] -= 100.0f * kLog2e;
      }
      // four independent chains for the row maximum and the row sum: with two softmax warps per scheduler there is little
      // else to hide the 4-cycle dependent-issue latency of a 49-long serial chain behind
      float m4[4] = {sv[0], sv[1], sv[2], sv[3]};
#pragma unroll
      for (int jj = 4; jj < 48; jj += 4) {
        m4[0] = fmaxf(m4[0], sv[jj]); m4[1] = fmaxf(m4[1], sv[jj + 1]); m4[2] = fmaxf(m4[2], sv[jj + 2]); m4[3] = fmaxf(m4[3], sv[jj + 3]);
      }
      const float mx = fmaxf(fmaxf(fmaxf(m4[0], sv[48]), m4[1]), fmaxf(m4[2], m4[3]));
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int jj = 0; jj < 52; jj += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float e = ex2f(sv[jj + u] - mx);
          s4[u] += e;
          sv[jj + u] = e;
        }
      }
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      QT(6);
      if (n >= 2) mbar_wait(bar(B_OFREE + g), ph ^ 1);       // the previous item's O tile (staged in sP) has been read by its TMA store
#pragma unroll
      for (int c = 0; c < 6; ++c) st_bf16x8(sP + sw128o(r, c), sv + 8 * c);
      {
        int4 pk;
        pk.x = pack_bf16(sv[48], sv[49]); pk.y = pack_bf16(sv[50], sv[51]); pk.z = 0; pk.w = 0;
        *reinterpret_cast<int4*>(sP + sw128o(r, 6)) = pk;
        *reinterpret_cast<int4*>(sP + sw128o(r, 7)) = make_int4(0, 0, 0, 0);
      }
      warp_arrive(B_P + g);
      QT(7);
      // ---- (c) O -> bf16 staging tile (sP is free: the P.V MMA has completed) -> the store warp's TMA store
      mbar_wait(bar(B_O + g), ph);
      tc_fence_after();
      QT(8);
      uint32_t o[32];
      tmem_ld32(tS + wloc * 32, o);
      tmem_ld_wait();
      {
        const float inv = valid ? __frcp_rn(sum) : 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) t[e] = __uint_as_float(o[8 * c + e]) * inv;
          st_bf16x8(sP + r * 64 + (((uint32_t)c ^ swz) << 4), t);
        }
        if (valid && p.lse != nullptr) p.lse[((size_t)win * p.nH + h) * QN + i] = (mx + __log2f(sum)) * 0.6931471805599453f;
      }
      QT(9);
      warp_arrive(B_OST + g);
      QT(10);
    }
    QT_PRINT;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace

// dynamic shared memory the kernel needs for (C, nH); 0 if the shape is not supported
static size_t attn_qkv_smem(int C, int nH, uint32_t* x_slot, uint32_t* w_head) {
  if (C % 32 != 0 || nH * QHD != C) return 0;
  const int nfull = C / 64, tail = (C % 64) ? 1 : 0;
  const uint32_t xs = nfull * kXFull + tail * kXTail, wh = nfull * kWFull + tail * kWTail;
  if (x_slot) *x_slot = xs;
  if (w_head) *w_head = wh;
  return (size_t)nH * wh + 2 * (size_t)xs + 2 * (size_t)(kQkvTileBytes + kPTileBytes) + (size_t)nH * QN * kRelLd * 4 + (size_t)3 * C * 4 + 1024;
}

int attn_qkv_supported(int C, int nH, int ws) {
  if (ws != 7) return 0;
  const size_t s = attn_qkv_smem(C, nH, nullptr, nullptr);
  return s != 0 && s <= kMaxDynSmem;
}

int attn_qkv_fwd(const swin_attn_qkv_args* a, cudaStream_t st) {
  SWIN_REQUIRE(a->ws == 7, "attn_qkv: window_size 7 only (got %d)", a->ws);
  SWIN_REQUIRE(a->B_ >= 0 && a->nH > 0, "attn_qkv: bad shape");
  SWIN_REQUIRE(a->x && a->wqkv && a->bias && a->out, "attn_qkv: null pointer");
  SWIN_REQUIRE(a->mask == nullptr || (a->nW > 0 && a->B_ % a->nW == 0), "attn_qkv: B_ must be a multiple of nW when a mask is given");
  const int C = a->nH * QHD;
  AttnQkvParams p;
  const size_t smem = attn_qkv_smem(C, a->nH, &p.x_slot_bytes, &p.w_head_bytes);
  SWIN_REQUIRE(smem != 0 && smem <= kMaxDynSmem, "attn_qkv: C = %d does not fit the resident-weight kernel (needs %zu bytes of shared memory)", C, smem);
  if (a->B_ == 0) return 0;
  p.B_ = a->B_; p.nH = a->nH; p.nW = a->nW > 0 ? a->nW : 1; p.C = C; p.ntiles = (a->B_ + 1) / 2;
  p.nfull = C / 64; p.tail = (C % 64) ? 1 : 0;
  p.scale = a->scale;
  p.rel_bias = a->bias; p.mask = a->mask; p.mask_nz = a->mask ? a->mask_nz : nullptr; p.bqkv = a->bqkv;
  p.canon_nwh = p.canon_nww = 0;
  if (a->mask && a->canon_nwh > 0 && a->canon_nww > 0) {
    SWIN_REQUIRE(a->canon_nwh * a->canon_nww == a->nW, "attn_qkv: canonical mask grid %d x %d does not match nW = %d", a->canon_nwh, a->canon_nww, a->nW);
    p.canon_nwh = a->canon_nwh; p.canon_nww = a->canon_nww;
  }
  p.lse = a->lse;
  p.qkv_out = (__nv_bfloat16*)a->qkv_out;
  const uint64_t rows = (uint64_t)a->B_ * QN;
  CUtensorMap tmX128, tmX64, tmW128, tmW64, tmOut;
  int rc;
  // 64-column (SW128) boxes for the full k-blocks, 32-column (SW64) boxes for the C % 64 == 32 tail
  if (p.nfull) {
    if ((rc = make_tmap_bf16_2d(&tmX128, a->x, (uint64_t)C, rows, (uint64_t)C * 2, 64, QN, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmW128, a->wqkv, (uint64_t)C, (uint64_t)3 * C, (uint64_t)C * 2, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  if ((rc = make_tmap_bf16_2d(&tmX64, a->x, (uint64_t)C, rows, (uint64_t)C * 2, 32, QN, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmW64, a->wqkv, (uint64_t)C, (uint64_t)3 * C, (uint64_t)C * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if (!p.nfull) { tmX128 = tmX64; tmW128 = tmW64; }
  if ((rc = make_tmap_bf16_2d(&tmOut, a->out, (uint64_t)C, rows, (uint64_t)C * 2, QHD, QN, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  rc = ensure_dyn_smem((const void*)attn_qkv_fwd_kernel, 0);       // the opt-in maximum minus the kernel's static shared memory
  if (rc) return rc;
  const int sms = persistent_sms();
  const int grid = p.ntiles < sms ? p.ntiles : sms;
  attn_qkv_fwd_kernel<<<grid, kQThreads, smem, st>>>(tmX128, tmX64, tmW128, tmW64, tmOut, p);
  SWIN_LAUNCH_CHECK();
  return 0;
}

}  // namespace swin
