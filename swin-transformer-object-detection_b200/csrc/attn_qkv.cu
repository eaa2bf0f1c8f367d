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
// Work item n = (window-pair tile k, head h), n = k * nH + h, walked in order by one CTA per SM.  Roles (320 threads):
//   warp 0      TMA producer: the weight slices once, then the X tile of every window pair into a 2-slot ring
//   warp 1      MMA issuer (one elected lane): per item  ACC = X W_h^T (128 x 96 x C)  ->  S = Q K^T (128 x 128 x 32)
//               ->  O = P [V0|V1] (128 x 64 x 64)
//   warps 2-5   softmax group A: items n even        warps 6-9   softmax group B: items n odd
//               per item: ACC (+bias) -> bf16 Q, K, V tiles in smem | S -> P (registers, thread = row) | O -> smem -> TMA store
// The two groups alternate items, so while one group converts / stores, the other runs its softmax and the tensor core works
// for both (the item itself is a serial chain; a second CTA per SM does not fit next to the resident weights).
// TMEM (512 columns): group g owns [256 g, 256 g + 256): ACC in [0, 96), S in [128, 256), O overlays S[0, 64).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace swin {
namespace {

constexpr int QHD = 32;            // head dim
constexpr int QN = 49;             // tokens per window (ws = 7)
constexpr int kQThreads = 320;
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
  int write_qkv;
  uint32_t x_slot_bytes, w_head_bytes;
};

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t sw128o(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ void st_bf16x8(uint8_t* dst, const float* v) {
  int4 pk;
  pk.x = pack_bf16(v[0], v[1]); pk.y = pack_bf16(v[2], v[3]); pk.z = pack_bf16(v[4], v[5]); pk.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<int4*>(dst) = pk;
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

// canonical SW-MSA mask of one row in closed form (see attn_tc.cu): bit j = 1 <=> mask[w][i][j] == -100
__device__ __forceinline__ unsigned long long canon_bits(int wi, int nwh, int nww, int i) {
  constexpr unsigned long long kRowHi = 0x1FFFFF0000000ULL, kColHi = 0x1C3870E1C3870ULL, kAll = 0x1FFFFFFFFFFFFULL;
  const int wh = wi / nww, ww = wi - wh * nww;
  unsigned long long m = 0ULL;
  if (wh == nwh - 1) m |= (i / 7 >= 4) ? (~kRowHi & kAll) : kRowHi;
  if (ww == nww - 1) m |= (i % 7 >= 4) ? (~kColHi & kAll) : kColHi;
  return m;
}

// barrier slots
enum { B_WFULL = 0, B_XFULL = 1, B_XEMPTY = 3, B_ACC = 5, B_QK = 7, B_S = 9, B_P = 11, B_O = 13, B_COUNT = 15 };

__global__ void __launch_bounds__(kQThreads, 1)
attn_qkv_fwd_kernel(const __grid_constant__ CUtensorMap tmX128, const __grid_constant__ CUtensorMap tmX64,
                    const __grid_constant__ CUtensorMap tmW128, const __grid_constant__ CUtensorMap tmW64,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmQKV, AttnQkvParams p) {
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
    for (int i = 0; i < B_COUNT; ++i) mbar_init(bar(i), 1);
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
        if (k >= 2) mbar_wait(bar(B_XEMPTY + slot), ((k >> 1) - 1) & 1);
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
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread walks the whole schedule)
    if (elect_one()) {
      constexpr uint32_t kHi64 = umma_desc_hi(512, (uint32_t)kSw64), kHi128 = umma_desc_hi(1024, (uint32_t)kSw128);
      const uint32_t idesc_qkv = umma_idesc_bf16(96, false, false);
      const uint32_t idesc_s = umma_idesc_bf16(128, false, false);
      const uint32_t idesc_o = umma_idesc_bf16(64, false, true);
      auto issue_qkv = [&](int n) {
        const int k = n / p.nH, h = n - k * p.nH, slot = k & 1, g = n & 1;
        if (h == 0) { mbar_wait(bar(B_XFULL + slot), (k >> 1) & 1); tc_fence_after(); }
        const uint32_t tAcc = tmem + g * 256;
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
        if (h == p.nH - 1) umma_commit(bar(B_XEMPTY + slot));       // the X slot is free once the last head's projection has run
      };
      auto issue_s = [&](int n) {
        const int g = n & 1;
        const uint32_t aQ = aG + g * (kQkvTileBytes + kPTileBytes), aK = aQ + 8192;
        const uint32_t qlo = umma_desc_lo(aQ, 16), klo = umma_desc_lo(aK, 16);
#pragma unroll
        for (uint32_t ks = 0; ks < 2; ++ks) umma_bf16(tmem + g * 256 + 128, umma_desc_join(kHi64, qlo + 2 * ks), umma_desc_join(kHi64, klo + 2 * ks), idesc_s, ks);
        umma_commit(bar(B_S + g));
      };
      auto issue_pv = [&](int n) {
        const int g = n & 1;
        const uint32_t aV = aG + g * (kQkvTileBytes + kPTileBytes) + 16384, aP = aV + 8192;
        const uint32_t plo = umma_desc_lo(aP, 16), vlo = umma_desc_lo(aV, 4096);
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk) umma_bf16(tmem + g * 256 + 128, umma_desc_join(kHi128, plo + 2 * kk), umma_desc_join(kHi64, vlo + 64 * kk), idesc_o, kk);
        umma_commit(bar(B_O + g));
      };
      const int nA = (nitems + 1) / 2, nB = nitems / 2;
      if (nitems > 0) {
        mbar_wait(bar(B_WFULL), 0);
        tc_fence_after();
        issue_qkv(0);
        if (nitems > 1) issue_qkv(1);
      }
      for (int j = 0; j < nA; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(bar(B_QK + 0), ph); tc_fence_after();             // group A: Q, K, V tiles of item 2j are in smem, ACC is free
        issue_s(2 * j);
        if (2 * j + 2 < nitems) issue_qkv(2 * j + 2);
        if (j > 0 && j - 1 < nB) { mbar_wait(bar(B_P + 1), (j - 1) & 1); tc_fence_after(); issue_pv(2 * j - 1); }
        if (j < nB) {
          mbar_wait(bar(B_QK + 1), ph); tc_fence_after();
          issue_s(2 * j + 1);
          if (2 * j + 3 < nitems) issue_qkv(2 * j + 3);
        }
        mbar_wait(bar(B_P + 0), ph); tc_fence_after();
        issue_pv(2 * j);
      }
      if (nB > 0 && nB == nA) { mbar_wait(bar(B_P + 1), (nB - 1) & 1); tc_fence_after(); issue_pv(2 * nB - 1); }
    }
    __syncwarp();
  } else {
    // ===================================================== softmax groups (thread = one row of the stacked 128-row tile)
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int r = q * 32 + lane, wloc = r >> 6, i = r & 63;
    const bool leader_warp = ((warp - 2) & 3) == 0;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tAcc = tmem + g * 256 + lane_off, tS = tAcc + 128;
    uint8_t* sQ = sG + g * (kQkvTileBytes + kPTileBytes);
    uint8_t* sP = sQ + kQkvTileBytes;
    const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
    const float sc2 = p.scale * kLog2e;
    const uint32_t swz = (uint32_t)((r >> 1) & 3);

    for (int n = g; n < nitems; n += 2) {
      const uint32_t ph = (uint32_t)(n >> 1) & 1;
      const int k = n / p.nH, h = n - k * p.nH;
      const int tile = blockIdx.x + k * G;
      const int win = 2 * tile + wloc;
      const bool valid = (i < QN) && (win < p.B_);
      // ---- (a) projection accumulator (+ bias) -> bf16 Q, K, V operand tiles
      mbar_wait(bar(B_ACC + g), ph);
      tc_fence_after();
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        uint32_t v[32];
        tmem_ld32(tAcc + part * 32, v);
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(sBq + part * p.C + h * QHD);
        uint8_t* trow = sQ + part * 8192 + r * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 b0 = b4[2 * c], b1 = b4[2 * c + 1];
          float t[8];
          t[0] = __uint_as_float(v[8 * c + 0]) + b0.x; t[1] = __uint_as_float(v[8 * c + 1]) + b0.y;
          t[2] = __uint_as_float(v[8 * c + 2]) + b0.z; t[3] = __uint_as_float(v[8 * c + 3]) + b0.w;
          t[4] = __uint_as_float(v[8 * c + 4]) + b1.x; t[5] = __uint_as_float(v[8 * c + 5]) + b1.y;
          t[6] = __uint_as_float(v[8 * c + 6]) + b1.z; t[7] = __uint_as_float(v[8 * c + 7]) + b1.w;
          st_bf16x8(trow + (((uint32_t)c ^ swz) << 4), t);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      if (leader_warp) {                             // the previous item's O tile (staged in sP) has been read out by its TMA store
        if (elect_one()) tma_store_wait_read<0>();
        __syncwarp();
      }
      group_sync(g);
      if (leader_warp) {
        if (elect_one()) {
          mbar_arrive(bar(B_QK + g));
          if (p.write_qkv) {                         // training: the stand-alone backward kernel reads qkv from HBM
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              if (2 * tile + w >= p.B_) continue;
#pragma unroll
              for (int part = 0; part < 3; ++part) tma_store_2d(&tmQKV, aQ + part * 8192 + w * 4096, part * p.C + h * QHD, (2 * tile + w) * QN);
            }
            tma_store_commit();
          }
        }
        __syncwarp();
      }
      // ---- (b) S -> P
      const float* mrow = nullptr;
      unsigned long long mb = 0ULL;
      if (p.mask != nullptr && valid) {
        const int mw = win % p.nW;
        if (p.canon_nwh > 0) mb = canon_bits(mw, p.canon_nwh, p.canon_nww, i);
        else if (p.mask_nz == nullptr || p.mask_nz[mw]) mrow = p.mask + ((size_t)mw * QN + i) * QN;
      }
      mbar_wait(bar(B_S + g), ph);
      tc_fence_after();
      uint32_t v[52];
      tmem_ld32(tS + wloc * 64, v);
      tmem_ld16(tS + wloc * 64 + 32, v + 32);
      tmem_ld4(tS + wloc * 64 + 48, v + 48);
      tmem_ld_wait();
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
      float mx = sv[0];
#pragma unroll
      for (int jj = 1; jj < QN; ++jj) mx = fmaxf(mx, sv[jj]);
      float sum = 0.f;
#pragma unroll
      for (int jj = 0; jj < 52; ++jj) {
        const float e = ex2f(sv[jj] - mx);
        sum += e;
        sv[jj] = e;
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) st_bf16x8(sP + sw128o(r, c), sv + 8 * c);
      {
        int4 pk;
        pk.x = pack_bf16(sv[48], sv[49]); pk.y = pack_bf16(sv[50], sv[51]); pk.z = 0; pk.w = 0;
        *reinterpret_cast<int4*>(sP + sw128o(r, 6)) = pk;
        *reinterpret_cast<int4*>(sP + sw128o(r, 7)) = make_int4(0, 0, 0, 0);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      group_sync(g);
      if (leader_warp) {
        if (elect_one()) mbar_arrive(bar(B_P + g));
        __syncwarp();
      }
      // ---- (c) O -> bf16 staging tile (sP is free: the P.V MMA has completed) -> TMA store
      mbar_wait(bar(B_O + g), ph);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(tS + wloc * 32, o);
      tmem_ld_wait();
      {
        const float inv = valid ? 1.0f / sum : 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) t[e] = __uint_as_float(o[8 * c + e]) * inv;
          st_bf16x8(sP + r * 64 + (((uint32_t)c ^ swz) << 4), t);
        }
        if (valid && p.lse != nullptr) p.lse[((size_t)win * p.nH + h) * QN + i] = (mx + log2f(sum)) * 0.6931471805599453f;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      if (leader_warp) {                             // this item's qkv stores have read the Q, K, V tiles (the next item rewrites them)
        if (elect_one()) tma_store_wait_read<0>();
        __syncwarp();
      }
      group_sync(g);
      if (leader_warp) {
        if (elect_one()) {
#pragma unroll
          for (int w = 0; w < 2; ++w)
            if (2 * tile + w < p.B_) tma_store_2d(&tmOut, aP + w * 4096, h * QHD, (2 * tile + w) * QN);
          tma_store_commit();
        }
        __syncwarp();
      }
    }
    if (leader_warp) {
      if (elect_one()) tma_store_wait_all<0>();
      __syncwarp();
    }
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
  p.write_qkv = a->qkv_out != nullptr;
  const uint64_t rows = (uint64_t)a->B_ * QN;
  CUtensorMap tmX128, tmX64, tmW128, tmW64, tmOut, tmQKV;
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
  tmQKV = tmOut;
  if (p.write_qkv)
    if ((rc = make_tmap_bf16_2d(&tmQKV, a->qkv_out, (uint64_t)3 * C, rows, (uint64_t)3 * C * 2, QHD, QN, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  rc = ensure_dyn_smem((const void*)attn_qkv_fwd_kernel, 0);       // the opt-in maximum minus the kernel's static shared memory
  if (rc) return rc;
  const int sms = persistent_sms();
  const int grid = p.ntiles < sms ? p.ntiles : sms;
  attn_qkv_fwd_kernel<<<grid, kQThreads, smem, st>>>(tmX128, tmX64, tmW128, tmW64, tmOut, tmQKV, p);
  SWIN_LAUNCH_CHECK();
  return 0;
}

}  // namespace swin
