// Fused QKV projection + window attention, forward (window 7, head_dim 32, bf16 operands, fp32 accumulation in TMEM).
//
//   REF = mmdet/models/backbones/swin_transformer.py:128-150:  qkv = x W^T + b ; S = scale q k^T + bias + mask ; P = softmax(S) ; O = P v
//
// One kernel replaces the qkv GEMM + the attention kernel: the LayerNorm'd window rows are read ONCE, Q / K / V of a head are
// produced by a tcgen05 GEMM against the weight slice of that head, converted to bf16 operand tiles in shared memory and
// consumed by the S / softmax / P.V pipeline without a round trip through HBM (training also TMA-stores the three tiles as
// the q, k, v rows the backward kernel reads).  HBM traffic per window pair is 2 x 49 x C bf16 in + the same out.
//
// Work item n = (window-pair tile k, head h), n = k * nH + h, walked in order by one CTA per SM.  Roles (416 threads):
//   warp 0      TMA producer: the X tile of every window pair (1- or 2-slot ring) and the weights -- either all heads' slices
//               once (RESIDENT: 3 C^2 bf16 fit next to the pipeline, C <= 128) or, per item, the k-blocks of that head's
//               96 x C slice through a ring of 12 KB slots (STREAM: C <= 384; the slices come from L2)
//   warp 1 / 2  attention MMA issuer of group A / B (one elected lane each): per item  S = Q K^T (128 x 128 x 32)  and
//               O = P [V0|V1] (128 x 64 x 64)
//   warp 3      store warp: TMA stores of the O tiles (and of the q, k, v tiles) of both groups, polled in arrival order
//   warp 12     projection MMA issuer: ACC_g = X W_h^T (128 x 96 x C) of every item, in item order, as soon as the group's
//               accumulator has been read out (i.e. under the PREVIOUS item's softmax).  It is the only consumer of the
//               weight ring, so a ring slot's full barrier is never tested more than one phase ahead.
//   warps 4-7   softmax group A: items n even        warps 8-11  softmax group B: items n odd
//               per item: ACC (+bias) -> bf16 Q, K, V tiles in smem | S -> P (registers, thread = row) | O -> bf16 staging tile
// The two groups alternate items, so while one group converts / stages, the other runs its softmax and the tensor core works
// for both.  Hand-overs are mbarriers with one arrival per warp (no CTA- or group-wide bar.sync inside the loop).
// Shared memory per group is 24 KB: the Q, K, V operand tiles; the P tile overlays Q + K (dead once S has been computed) and
// the O staging tile overlays V (dead once P.V has completed).
// TMEM (512 columns): group g owns [256 g, 256 g + 256): ACC in [0, 96), S in [128, 256), O overlays S[0, 64).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace swin {
namespace {

constexpr int QHD = 32;            // head dim
constexpr int QN = 49;             // tokens per window (ws = 7)
constexpr int kQThreads = 416;
constexpr int kRelLd = 52;         // key columns per row handled by the softmax: 49 + 3 pad columns (bias kNegBigQ -> P = 0)
constexpr int kRelI = 64;          // query-row pitch of the transposed bias table (workspace): (nH, 52, 64) floats
constexpr float kNegBigQ = -1.0e30f;
constexpr float kLog2eQ = 1.4426950408889634f;
constexpr size_t kMaxDynSmem = 227 * 1024 - 4096;     // opt-in limit minus head-room for the kernel's static shared memory
constexpr uint32_t kGroupBytes = 3 * 8192;      // Q, K, V operand tiles of one group: 128 rows x 32 bf16 each (SW64), window 1 at +4096
constexpr uint32_t kXFull = 16384, kXTail = 8192;      // X k-blocks: 128 rows x 64 (SW128) / x 32 (SW64) bf16
constexpr uint32_t kWFull = 12288, kWTail = 6144;      // W_h k-blocks: 96 rows x 64 (SW128) / x 32 (SW64) bf16
constexpr int kMaxWSlots = 12;

struct AttnQkvParams {
  int B_, nH, nW, C, ntiles;
  int nfull, tail, nkb;            // C = 64 * nfull + 32 * tail ; nkb = nfull + tail k-blocks per projection
  int stream_w, wslots, xslots;
  int canon_nwh, canon_nww;
  int want_qkv;
  float scale;
  const float* rel_t;              // workspace: (nH, 52 key columns, 64 query rows), pre-scaled by log2(e), pad columns kNegBigQ:
                                   // lane = query row reads one coalesced line per key column (L1-resident, no shared memory)
  const float* mask; const int* mask_nz; const float* bqkv;
  float* lse;
  uint32_t x_slot_bytes, w_head_bytes;
};

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t sw128o(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

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

// barrier slots (pairs are indexed by X slot or by group)
enum { B_XFULL = 0, B_XEMPTY = 2, B_ACC = 4, B_QK = 6, B_S = 8, B_P = 10, B_O = 12, B_OST = 14, B_OFREE = 16, B_QKVFREE = 18,
       B_WRES = 20, B_WFULL = 21, B_WEMPTY = B_WFULL + kMaxWSlots, B_COUNT = B_WEMPTY + kMaxWSlots };

// polling loops of single threads that watch SEVERAL barriers: bounded like mbar_wait (a protocol bug traps instead of hanging)
struct SpinGuard {
  long long t0 = 0; int spins = 0;
  __device__ __forceinline__ void tick() {
    if (spins == 0) t0 = clock64();
    if ((++spins & 4095) == 0 && clock64() - t0 > 4000000000LL) { atomicExch(&g_watchdog_flag, 1); __trap(); }
  }
};
// Canonical SW-MSA mask (REF:370-389) in closed form, window 7 / shift 3 (see attn_tc.cu: canon_mask_pen): a key column's
// (rowhi, colhi) is a compile-time class in the unrolled loops, a row needs four constants pen[rowhi][colhi] in {0, -100 log2 e},
// and they fold into the "minus the row maximum" constant of the exponent -- the mask costs no per-element instruction.
struct MaskPenQ { float c[2][2]; };
__device__ __forceinline__ MaskPenQ canon_pen(bool active, int wi, int nwh, int nww, int i) {
  MaskPenQ m;
  bool R = false, Cw = false;
  if (active) {
    const int wh = wi / nww, ww = wi - wh * nww;
    R = wh == nwh - 1; Cw = ww == nww - 1;
  }
  const int ii = i < QN ? i : QN - 1;
  const bool ri = ii / 7 >= 4, ci = ii % 7 >= 4;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) m.c[a][b] = ((R && ((a != 0) != ri)) || (Cw && ((b != 0) != ci))) ? -100.0f * kLog2eQ : 0.f;
  return m;
}
#define QRH(j) (((j) / 7) >= 4 ? 1 : 0)
#define QCH(j) (((j) % 7) >= 4 ? 1 : 0)

__global__ void __launch_bounds__(kQThreads, 1)
attn_qkv_fwd_kernel(const __grid_constant__ CUtensorMap tmX128, const __grid_constant__ CUtensorMap tmX64,
                    const __grid_constant__ CUtensorMap tmW128, const __grid_constant__ CUtensorMap tmW64,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmQkv, AttnQkvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[B_COUNT];
  __shared__ uint32_t tmem_slot;
  uint8_t* sbase = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = sbase;                                              // RESIDENT: nH x w_head_bytes ; STREAM: wslots x kWFull
  const uint32_t w_bytes = p.stream_w ? (uint32_t)p.wslots * kWFull : (uint32_t)p.nH * p.w_head_bytes;
  uint8_t* sX = sW + w_bytes;                                       // xslots x x_slot_bytes
  uint8_t* sG = sX + (uint32_t)p.xslots * p.x_slot_bytes;           // 2 groups x {Q, K, V tiles}
  float* sBq = reinterpret_cast<float*>(sG + 2u * kGroupBytes);     // 3C qkv bias (zeros if the Linear has none)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };

  // ---- prologue: zero the X ring (its pad rows 49..63 are never written by TMA), stage the qkv bias
  for (uint32_t o = tid * 16; o < (uint32_t)p.xslots * p.x_slot_bytes; o += kQThreads * 16) *reinterpret_cast<int4*>(sX + o) = make_int4(0, 0, 0, 0);
  for (int e = tid; e < 3 * p.C; e += kQThreads) sBq[e] = p.bqkv != nullptr ? p.bqkv[e] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i) {
      const bool per_warp = (i >= B_QK && i < B_QK + 2) || (i >= B_P && i < B_P + 2) || (i >= B_OST && i < B_OST + 2);   // one arrival per group warp
      mbar_init(bar(i), per_warp ? 4 : 1);
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
    // ===================================================== TMA producer: two cursors (X tiles, W k-blocks), neither waits for the other
    if (elect_one()) {
      if (!p.stream_w) {
        mbar_expect_tx(bar(B_WRES), (uint32_t)(3 * p.C * p.C * 2));
        for (int h = 0; h < p.nH; ++h) {
          const uint32_t wb = aW + (uint32_t)h * p.w_head_bytes;
          for (int kb = 0; kb < p.nfull; ++kb)
            for (int part = 0; part < 3; ++part) tma_load_2d(wb + kb * kWFull + part * 4096, &tmW128, bar(B_WRES), kb * 64, part * p.C + h * QHD);
          if (p.tail)
            for (int part = 0; part < 3; ++part) tma_load_2d(wb + p.nfull * kWFull + part * 2048, &tmW64, bar(B_WRES), p.nfull * 64, part * p.C + h * QHD);
        }
      }
      int xk = 0;                                    // next X tile to request
      int wn = 0, wkb = 0, wslot = 0, wuse = 0;      // next W k-block: item, k-block, ring slot and how often that slot has been used
      const bool wstream = p.stream_w != 0;
      SpinGuard sg;
      while (xk < nk || (wstream && wn < nitems)) {
        bool progressed = false;
        if (xk < nk) {
          const int slot = xk % p.xslots, use = xk / p.xslots;
          if (use == 0 || mbar_test_wait(bar(B_XEMPTY + slot), (uint32_t)(use - 1) & 1)) {
            const uint32_t xb = aX + (uint32_t)slot * p.x_slot_bytes, fb = bar(B_XFULL + slot);
            mbar_expect_tx(fb, (uint32_t)(2 * QN * p.C * 2));
            const int tile = blockIdx.x + xk * G;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              const int row0 = (2 * tile + w) * QN;             // rows past the tensor (odd window count) arrive as zeros
              for (int kb = 0; kb < p.nfull; ++kb) tma_load_2d(xb + kb * kXFull + w * 8192, &tmX128, fb, kb * 64, row0);
              if (p.tail) tma_load_2d(xb + p.nfull * kXFull + w * 4096, &tmX64, fb, p.nfull * 64, row0);
            }
            ++xk;
            progressed = true;
          }
        }
        if (wstream && wn < nitems) {
          if (wuse == 0 || mbar_test_wait(bar(B_WEMPTY + wslot), (uint32_t)(wuse - 1) & 1)) {
            const int h = wn % p.nH;
            const uint32_t wb = aW + (uint32_t)wslot * kWFull, fb = bar(B_WFULL + wslot);
            if (wkb < p.nfull) {
              mbar_expect_tx(fb, kWFull);
#pragma unroll
              for (int part = 0; part < 3; ++part) tma_load_2d(wb + part * 4096, &tmW128, fb, wkb * 64, part * p.C + h * QHD);
            } else {
              mbar_expect_tx(fb, kWTail);
#pragma unroll
              for (int part = 0; part < 3; ++part) tma_load_2d(wb + part * 2048, &tmW64, fb, p.nfull * 64, part * p.C + h * QHD);
            }
            if (++wkb == p.nkb) { wkb = 0; ++wn; }
            if (++wslot == p.wslots) { wslot = 0; ++wuse; }
            progressed = true;
          }
        }
        if (!progressed) { __nanosleep(32); sg.tick(); }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2) {
    // ===================================================== attention MMA issuer of group g
    const int g = warp - 1;
    if (elect_one()) {
      constexpr uint32_t kHi64 = umma_desc_hi(512, (uint32_t)kSw64), kHi128 = umma_desc_hi(1024, (uint32_t)kSw128);
      const uint32_t idesc_s = umma_idesc_bf16(128, false, false);
      const uint32_t idesc_o = umma_idesc_bf16(64, false, true);
      const uint32_t tS = tmem + g * 256 + 128;
      const uint32_t aQ = aG + g * kGroupBytes, aK = aQ + 8192, aV = aQ + 16384, aP = aQ;
      const uint32_t qlo = umma_desc_lo(aQ, 16), klo = umma_desc_lo(aK, 16), plo = umma_desc_lo(aP, 16), vlo = umma_desc_lo(aV, 4096);
      for (int n = g; n < nitems; n += 2) {
        const uint32_t ph = (uint32_t)(n >> 1) & 1;
        mbar_wait(bar(B_QK + g), ph);                // Q, K, V tiles of item n are in smem
        tc_fence_after();
#pragma unroll
        for (uint32_t ks = 0; ks < 2; ++ks) umma_bf16(tS, umma_desc_join(kHi64, qlo + 2 * ks), umma_desc_join(kHi64, klo + 2 * ks), idesc_s, ks);
        umma_commit(bar(B_S + g));
        mbar_wait(bar(B_P + g), ph);                 // P tile written (over Q + K), S read out
        tc_fence_after();
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk) umma_bf16(tS, umma_desc_join(kHi128, plo + 2 * kk), umma_desc_join(kHi64, vlo + 64 * kk), idesc_o, kk);
        umma_commit(bar(B_O + g));
      }
    }
    __syncwarp();
  } else if (warp == 12) {
    // ===================================================== projection MMA issuer: every item, in order
    if (elect_one()) {
      constexpr uint32_t kHi64 = umma_desc_hi(512, (uint32_t)kSw64), kHi128 = umma_desc_hi(1024, (uint32_t)kSw128);
      const uint32_t idesc_qkv = umma_idesc_bf16(96, false, false);
      const bool wstream = p.stream_w != 0;
      if (!wstream && nitems > 0) mbar_wait(bar(B_WRES), 0);
      int k = 0, h = 0;                              // item n = (k, h)
      int wslot = 0;
      uint32_t wpar = 0;                             // ring cursor: slot and the parity of its current use
      for (int n = 0; n < nitems; ++n) {
        const int g = n & 1, slot = k % p.xslots;
        const uint32_t tAcc = tmem + g * 256;
        if (h == 0) mbar_wait(bar(B_XFULL + slot), (uint32_t)(k / p.xslots) & 1);
        if (n >= 2) mbar_wait(bar(B_QK + g), (uint32_t)((n - 2) >> 1) & 1);       // the group has read ACC of its previous item
        tc_fence_after();
        const uint32_t xb0 = aX + (uint32_t)slot * p.x_slot_bytes;
        for (int kb = 0; kb < p.nkb; ++kb) {
          uint32_t wb;
          if (wstream) {
            mbar_wait(bar(B_WFULL + wslot), wpar);
            tc_fence_after();
            wb = aW + (uint32_t)wslot * kWFull;
          } else {
            wb = aW + (uint32_t)h * p.w_head_bytes + (uint32_t)kb * kWFull;
          }
          const uint32_t xlo = umma_desc_lo(xb0 + (uint32_t)kb * kXFull, 16), wlo = umma_desc_lo(wb, 16);
          if (kb < p.nfull) {
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks) umma_bf16(tAcc, umma_desc_join(kHi128, xlo + 2 * ks), umma_desc_join(kHi128, wlo + 2 * ks), idesc_qkv, (kb | ks) != 0);
          } else {
#pragma unroll
            for (uint32_t ks = 0; ks < 2; ++ks) umma_bf16(tAcc, umma_desc_join(kHi64, xlo + 2 * ks), umma_desc_join(kHi64, wlo + 2 * ks), idesc_qkv, (kb | ks) != 0);
          }
          if (wstream) {
            umma_commit(bar(B_WEMPTY + wslot));
            if (++wslot == p.wslots) { wslot = 0; wpar ^= 1; }
          }
        }
        umma_commit(bar(B_ACC + g));
        if (++h == p.nH) { umma_commit(bar(B_XEMPTY + slot)); h = 0; ++k; }      // every head's projection of this tile has been issued
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================================================== store warp: O tiles (and q, k, v tiles) of both groups, in arrival order
    if (elect_one()) {
      int nq = p.want_qkv ? 0 : nitems, no = 0;      // next item whose q/k/v tiles / whose O tile is still to be stored
      SpinGuard sg;
      while (no < nitems) {
        bool progressed = false;
        if (nq < nitems && mbar_test_wait(bar(B_QK + (nq & 1)), (uint32_t)(nq >> 1) & 1)) {
          const int g = nq & 1, k = nq / p.nH, h = nq - k * p.nH, tile = blockIdx.x + k * G;
          const uint32_t aQ = aG + g * kGroupBytes;
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            if (2 * tile + w >= p.B_) continue;
#pragma unroll
            for (int part = 0; part < 3; ++part) tma_store_2d(&tmQkv, aQ + part * 8192 + w * 4096, part * p.C + h * QHD, (2 * tile + w) * QN);
          }
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(bar(B_QKVFREE + g));
          ++nq; progressed = true;
        }
        if (mbar_test_wait(bar(B_OST + (no & 1)), (uint32_t)(no >> 1) & 1)) {
          const int g = no & 1, k = no / p.nH, h = no - k * p.nH, tile = blockIdx.x + k * G;
          const uint32_t aV = aG + g * kGroupBytes + 16384;
#pragma unroll
          for (int w = 0; w < 2; ++w)
            if (2 * tile + w < p.B_) tma_store_2d(&tmOut, aV + w * 4096, h * QHD, (2 * tile + w) * QN);
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(bar(B_OFREE + g));
          ++no; progressed = true;
        }
        if (!progressed) { __nanosleep(32); sg.tick(); }
      }
      tma_store_wait_all<0>();
    }
    __syncwarp();
  } else {
    // ===================================================== softmax groups (warps 4-11; thread = one row of the stacked 128-row tile)
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int r = q * 32 + lane, wloc = r >> 6, i = r & 63;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tAcc = tmem + g * 256 + lane_off, tS = tAcc + 128;
    uint8_t* sQ = sG + g * kGroupBytes;
    uint8_t* sP = sQ;                                // P overlays Q + K
    uint8_t* sO = sQ + 16384;                        // O staging overlays V
    const f32x2 sc2 = pk2(p.scale * kLog2eQ, p.scale * kLog2eQ);
    const uint32_t swz = (uint32_t)((r >> 1) & 3);
    // hand-over to the MMA / store warps: every thread has fenced its smem writes (generic -> async proxy) and its TMEM reads,
    // then one lane per warp arrives (the barriers expect 4 arrivals)
    auto warp_arrive = [&](int b) {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(b));
    };
    // one 32-column part of the accumulator (+ bias) -> bf16 operand tile rows
    auto convert_part = [&](const uint32_t* vv, int part, int h) {
      const float4* b4 = reinterpret_cast<const float4*>(sBq + part * p.C + h * QHD);
      uint8_t* trow = sQ + part * 8192 + r * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 b0 = b4[2 * c], b1 = b4[2 * c + 1];
        float x0, x1, x2, x3, x4, x5, x6, x7;
        unpk2(add2(pk2u(vv[8 * c + 0], vv[8 * c + 1]), pk2(b0.x, b0.y)), x0, x1);
        unpk2(add2(pk2u(vv[8 * c + 2], vv[8 * c + 3]), pk2(b0.z, b0.w)), x2, x3);
        unpk2(add2(pk2u(vv[8 * c + 4], vv[8 * c + 5]), pk2(b1.x, b1.y)), x4, x5);
        unpk2(add2(pk2u(vv[8 * c + 6], vv[8 * c + 7]), pk2(b1.z, b1.w)), x6, x7);
        int4 pk;
        pk.x = pack_bf16(x0, x1); pk.y = pack_bf16(x2, x3); pk.z = pack_bf16(x4, x5); pk.w = pack_bf16(x6, x7);
        *reinterpret_cast<int4*>(trow + (((uint32_t)c ^ swz) << 4)) = pk;
      }
    };

    QT_DECL
    int mask_k = -1;
    const float* mrow = nullptr;
    MaskPenQ pen = canon_pen(false, 0, 1, 1, i);
    for (int n = g; n < nitems; n += 2) {
      const uint32_t ph = (uint32_t)(n >> 1) & 1;
      QT_ITEM
      const int k = n / p.nH, h = n - k * p.nH;
      const int tile = blockIdx.x + k * G;
      const int win = 2 * tile + wloc;
      const bool valid = (i < QN) && (win < p.B_);
      // ---- (a) projection accumulator (+ bias) -> bf16 Q, K, V operand tiles
      mbar_wait(bar(B_ACC + g), ph);
      tc_fence_after();
      QT(0);
      {
        // 32 accumulator columns at a time, the next part's TMEM load in flight while this one is converted
        uint32_t va[32], vb[32];
        tmem_ld32(tAcc, va);
        tmem_ld_wait();
        tmem_ld32(tAcc + 32, vb);
        convert_part(va, 0, h);                      // Q, K: the P tile of item n - 2 that overlays them died with its P.V MMA
        tmem_ld_wait();
        tmem_ld32(tAcc + 64, va);
        convert_part(vb, 1, h);
        tmem_ld_wait();
        if (n >= 2) mbar_wait(bar(B_OFREE + g), ph ^ 1);     // V: item n - 2's O tile (staged there) has been read by its TMA store
        convert_part(va, 2, h);
      }
      QT(1);
      warp_arrive(B_QK + g);
      QT(2);
      // ---- (b) S -> P
      if (k != mask_k) {                              // the row's mask depends on the window only: once per tile, not per head
        mask_k = k; mrow = nullptr;
        bool canon = false;
        if (p.mask != nullptr && valid) {
          if (p.canon_nwh > 0) canon = true;
          else if (p.mask_nz == nullptr || p.mask_nz[win % p.nW]) mrow = p.mask + ((size_t)(win % p.nW) * QN + i) * QN;
        }
        pen = canon_pen(canon, canon ? win % p.nW : 0, p.canon_nwh, p.canon_nww, i);
      }
      // this row's bias values: 52 coalesced loads (one line per key column across the warp), issued before the wait for S
      float bv[52];
      {
        const float* bt = p.rel_t + (size_t)h * (kRelLd * kRelI) + i;
#pragma unroll
        for (int c = 0; c < 52; ++c) bv[c] = __ldg(bt + c * kRelI);
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
      f32x2 s2[26];
#pragma unroll
      for (int c = 0; c < 26; ++c) s2[c] = fma2(pk2u(v[2 * c], v[2 * c + 1]), sc2, pk2(bv[2 * c], bv[2 * c + 1]));
      float sv[52];
#pragma unroll
      for (int c = 0; c < 26; ++c) unpk2(s2[c], sv[2 * c], sv[2 * c + 1]);
      if (mrow != nullptr) {
#pragma unroll
        for (int jj = 0; jj < QN; ++jj) sv[jj] = fmaf(__ldg(mrow + jj), kLog2eQ, sv[jj]);
      }
      // row maximum of the MASKED logits from the four class maxima; exponent offset per class = penalty - maximum
      float cm[2][2] = {{sv[0], sv[4]}, {sv[28], sv[32]}};
#pragma unroll
      for (int jj = 1; jj < QN; ++jj) cm[QRH(jj)][QCH(jj)] = fmaxf(cm[QRH(jj)][QCH(jj)], sv[jj]);
      const float mx = fmaxf(fmaxf(cm[0][0] + pen.c[0][0], cm[0][1] + pen.c[0][1]), fmaxf(cm[1][0] + pen.c[1][0], cm[1][1] + pen.c[1][1]));
      const float off[2][2] = {{pen.c[0][0] - mx, pen.c[0][1] - mx}, {pen.c[1][0] - mx, pen.c[1][1] - mx}};
      f32x2 acc2[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
      uint32_t pb[26];                                // P row as packed bf16 pairs
#pragma unroll
      for (int c = 0; c < 26; ++c) {
        const float ea = ex2f(sv[2 * c] + off[QRH(2 * c)][QCH(2 * c)]), eb = ex2f(sv[2 * c + 1] + off[QRH(2 * c + 1)][QCH(2 * c + 1)]);
        acc2[c & 1] = add2(acc2[c & 1], pk2(ea, eb));
        pb[c] = pack_bf16(ea, eb);
      }
      float sum;
      {
        float a, b;
        unpk2(add2(acc2[0], acc2[1]), a, b);
        sum = a + b;
      }
      QT(6);
      if (p.want_qkv) mbar_wait(bar(B_QKVFREE + g), ph);     // the TMA stores of this item's q, k, v tiles have read them
#pragma unroll
      for (int c = 0; c < 6; ++c) *reinterpret_cast<int4*>(sP + sw128o(r, c)) = make_int4((int)pb[4 * c], (int)pb[4 * c + 1], (int)pb[4 * c + 2], (int)pb[4 * c + 3]);
      *reinterpret_cast<int4*>(sP + sw128o(r, 6)) = make_int4((int)pb[24], (int)pb[25], 0, 0);
      *reinterpret_cast<int4*>(sP + sw128o(r, 7)) = make_int4(0, 0, 0, 0);
      warp_arrive(B_P + g);
      QT(7);
      // ---- (c) O -> bf16 staging tile (the V tile is dead: the P.V MMA has completed) -> the store warp's TMA store
      mbar_wait(bar(B_O + g), ph);
      tc_fence_after();
      QT(8);
      uint32_t o[32];
      tmem_ld32(tS + wloc * 32, o);
      tmem_ld_wait();
      {
        const float inv = valid ? __frcp_rn(sum) : 0.f;
        const f32x2 inv2 = pk2(inv, inv);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t0, t1, t2, t3, t4, t5, t6, t7;
          unpk2(mul2(pk2u(o[8 * c + 0], o[8 * c + 1]), inv2), t0, t1);
          unpk2(mul2(pk2u(o[8 * c + 2], o[8 * c + 3]), inv2), t2, t3);
          unpk2(mul2(pk2u(o[8 * c + 4], o[8 * c + 5]), inv2), t4, t5);
          unpk2(mul2(pk2u(o[8 * c + 6], o[8 * c + 7]), inv2), t6, t7);
          int4 pk;
          pk.x = pack_bf16(t0, t1); pk.y = pack_bf16(t2, t3); pk.z = pack_bf16(t4, t5); pk.w = pack_bf16(t6, t7);
          *reinterpret_cast<int4*>(sO + r * 64 + (((uint32_t)c ^ swz) << 4)) = pk;
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

// (nH, 49, 49) fp32 dense bias -> transposed table (nH, 52, 64): [h][j][i] = bias[h][min(i, 48)][j] * log2(e), kNegBigQ for j >= 49
__global__ void rel_t_kernel(const float* __restrict__ bias, float* __restrict__ out, int nH) {
  const int total = nH * kRelLd * kRelI;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int hh = e / (kRelLd * kRelI), rem = e - hh * (kRelLd * kRelI), j = rem / kRelI, i = rem - j * kRelI;
    out[e] = j < QN ? bias[((size_t)hh * QN + (i < QN ? i : QN - 1)) * QN + j] * kLog2eQ : kNegBigQ;
  }
}

struct QkvPlan {
  int ok, stream_w, wslots, xslots, nfull, tail;
  uint32_t x_slot, w_head;
  size_t smem;
};

// shared-memory plan for (C, nH): RESIDENT weights if they fit, else the k-block ring; ok = 0 if neither does
QkvPlan attn_qkv_plan(int C, int nH) {
  QkvPlan pl = {};
  if (C % 32 != 0 || nH * QHD != C || C < 32) return pl;
  pl.nfull = C / 64; pl.tail = (C % 64) ? 1 : 0;
  pl.x_slot = pl.nfull * kXFull + pl.tail * kXTail;
  pl.w_head = pl.nfull * kWFull + pl.tail * kWTail;
  const size_t fixed = 2 * (size_t)kGroupBytes + (size_t)3 * C * 4 + 1024;
  const size_t resident = fixed + (size_t)nH * pl.w_head + 2 * (size_t)pl.x_slot;
  if (resident <= kMaxDynSmem) {
    pl.ok = 1; pl.stream_w = 0; pl.wslots = 0; pl.xslots = 2; pl.smem = resident;
    return pl;
  }
  // STREAM: two X slots if at least 6 ring slots remain next to them, else one
  for (int xs = 2; xs >= 1; --xs) {
    const size_t base = fixed + (size_t)xs * pl.x_slot;
    if (base >= kMaxDynSmem) continue;
    int slots = (int)((kMaxDynSmem - base) / kWFull);
    if (slots > kMaxWSlots) slots = kMaxWSlots;
    if (slots >= (xs == 2 ? 6 : 4)) {
      pl.ok = 1; pl.stream_w = 1; pl.wslots = slots; pl.xslots = xs; pl.smem = base + (size_t)slots * kWFull;
      return pl;
    }
  }
  return pl;
}

}  // namespace

int attn_qkv_supported(int C, int nH, int ws) {
  if (ws != 7) return 0;
  return attn_qkv_plan(C, nH).ok;
}

long long attn_qkv_workspace_bytes(int C, int nH, int ws) {
  if (ws != 7) return 0;
  const QkvPlan pl = attn_qkv_plan(C, nH);
  return pl.ok ? (long long)nH * kRelLd * kRelI * 4 : 0;
}

int attn_qkv_fwd(const swin_attn_qkv_args* a, cudaStream_t st) {
  SWIN_REQUIRE(a->ws == 7, "attn_qkv: window_size 7 only (got %d)", a->ws);
  SWIN_REQUIRE(a->B_ >= 0 && a->nH > 0, "attn_qkv: bad shape");
  SWIN_REQUIRE(a->x && a->wqkv && a->bias && a->out, "attn_qkv: null pointer");
  SWIN_REQUIRE(a->mask == nullptr || (a->nW > 0 && a->B_ % a->nW == 0), "attn_qkv: B_ must be a multiple of nW when a mask is given");
  const int C = a->nH * QHD;
  const QkvPlan pl = attn_qkv_plan(C, a->nH);
  SWIN_REQUIRE(pl.ok, "attn_qkv: C = %d does not fit the fused kernel (window-pair tile + weight ring exceed shared memory)", C);
  AttnQkvParams p;
  {
    const size_t need = (size_t)a->nH * kRelLd * kRelI * 4;
    SWIN_REQUIRE(a->workspace != nullptr && a->workspace_bytes >= (long long)need && aligned16(a->workspace),
                 "attn_qkv: needs a %zu-byte 16-byte-aligned workspace (swin_window_attn_qkv_workspace)", need);
    p.rel_t = (const float*)a->workspace;
  }
  if (a->B_ == 0) return 0;
  p.B_ = a->B_; p.nH = a->nH; p.nW = a->nW > 0 ? a->nW : 1; p.C = C; p.ntiles = (a->B_ + 1) / 2;
  p.nfull = pl.nfull; p.tail = pl.tail; p.nkb = pl.nfull + pl.tail;
  p.stream_w = pl.stream_w; p.wslots = pl.wslots; p.xslots = pl.xslots;
  p.x_slot_bytes = pl.x_slot; p.w_head_bytes = pl.w_head;
  p.scale = a->scale;
  p.mask = a->mask; p.mask_nz = a->mask ? a->mask_nz : nullptr; p.bqkv = a->bqkv;
  p.canon_nwh = p.canon_nww = 0;
  if (a->mask && a->canon_nwh > 0 && a->canon_nww > 0) {
    SWIN_REQUIRE(a->canon_nwh * a->canon_nww == a->nW, "attn_qkv: canonical mask grid %d x %d does not match nW = %d", a->canon_nwh, a->canon_nww, a->nW);
    p.canon_nwh = a->canon_nwh; p.canon_nww = a->canon_nww;
  }
  p.lse = a->lse;
  p.want_qkv = a->qkv_out != nullptr;
  const uint64_t rows = (uint64_t)a->B_ * QN;
  CUtensorMap tmX128, tmX64, tmW128, tmW64, tmOut, tmQkv;
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
  tmQkv = tmOut;
  if (p.want_qkv) {
    SWIN_REQUIRE(aligned16(a->qkv_out), "attn_qkv: qkv_out alignment");
    if ((rc = make_tmap_bf16_2d(&tmQkv, a->qkv_out, (uint64_t)3 * C, rows, (uint64_t)3 * C * 2, QHD, QN, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  }
  rc = ensure_dyn_smem((const void*)attn_qkv_fwd_kernel, 0);       // the opt-in maximum minus the kernel's static shared memory
  if (rc) return rc;
  rel_t_kernel<<<(a->nH * kRelLd * kRelI + 255) / 256, 256, 0, st>>>(a->bias, (float*)a->workspace, a->nH);
  SWIN_LAUNCH_CHECK();
  const int sms = persistent_sms();
  const int grid = p.ntiles < sms ? p.ntiles : sms;
  attn_qkv_fwd_kernel<<<grid, kQThreads, pl.smem, st>>>(tmX128, tmX64, tmW128, tmW64, tmOut, tmQkv, p);
  SWIN_LAUNCH_CHECK();
  return 0;
}

}  // namespace swin
