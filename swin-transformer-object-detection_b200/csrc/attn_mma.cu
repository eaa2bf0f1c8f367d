// Window attention core for 144-token windows (window 12: the 384-pixel Swin-B/L configs), bf16 storage, forward + backward.
//
//   REF = mmdet/models/backbones/swin_transformer.py:132-150:  S = scale q k^T + bias + mask ; P = softmax(S) ; O = P v
//
// The tcgen05 kernels (attn_tc.cu) are specialised for 49-token windows stacked two per 128-row tile; a 144-token window needs
// another tile scheme.  The op is HBM-bound (24.5 flop/B), so this kernel uses the warp-level tensor-core path
// (mma.sync.m16n8k16, bf16 x bf16 -> fp32) FlashAttention-2 style: one CTA per (window, head), nine warps, warp = 16 query
// rows against all 144 keys held as register fragments; q / k / v (/ dO / O) tiles arrive by 16-byte cp.async into a
// two-stage shared-memory ring (next item in flight while this one computes); softmax statistics in fp32.
// Backward: phase 1 (warp = 16 query rows) recomputes P from the saved LSE in three 48-key chunks, forms
// dS = P (dP - D) with D = rowsum(dO . O), accumulates dQ and writes P / scale dS (bf16) to shared memory; phase 2
// (warp = 16 keys) reads them transposed (ldmatrix.trans) for dV = P^T dO and dK = dS^T Q.  dBias lives in registers across
// the items of a CTA (a CTA stays on one head) and is added to global memory once.
// 2.3-7 TF/s fp32 FFMA kernels (attn_simt.cu) remain the fp32 parity path and serve every other window size.
#include "common.cuh"

namespace swin {
namespace {

constexpr int MHD = 32;
constexpr int kPitch = 80;          // bytes per row of a [rows][32] bf16 tile in shared memory: 64 data + 16 pad (ldmatrix conflict-free)
constexpr float kL2e = 1.4426950408889634f;

struct AttnMmaParams {
  int B_, nH, nW, C, per_head;
  float scale;
  const __nv_bfloat16* qkv; const float* bias; const float* mask; const int* mask_nz;
  __nv_bfloat16* out; float* lse;
  const __nv_bfloat16* dout; __nv_bfloat16* dqkv; float* dbias;
};

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D (16x8 fp32) += A (16x16 bf16, row) * B (16x8 bf16, col)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// [rows][32] bf16 tile: global (row stride `ld` elements) -> shared (pitch kPitch), 16 bytes per cp.async
template <int ROWS>
__device__ __forceinline__ void load_tile(uint8_t* sdst, const __nv_bfloat16* gsrc, int ld) {
  for (int idx = threadIdx.x; idx < ROWS * 4; idx += blockDim.x) {
    const int row = idx >> 2, c = idx & 3;
    cp_async16(sm_u32(sdst + row * kPitch + c * 16), gsrc + (size_t)row * ld + c * 8);
  }
}
// A fragments (two k-steps over d = 32) of the 16 rows starting at row0 of a [rows][32] tile
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[2][4], const uint8_t* tile, int row0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    ldsm_x4(a[ks], sm_u32(tile + (row0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kPitch + (ks * 16 + 8 * (lane >> 4)) * 2));
}
// acc[2 nt2], acc[2 nt2 + 1] (two 8-key n-tiles) += A(16 x 32) * T[keys 16 nt2 .. +15][0..31]^T     (T rows = n, T cols = k)
__device__ __forceinline__ void mma_nk(float (&c0)[4], float (&c1)[4], const uint32_t (&a)[2][4], const uint8_t* tile, int nt2, int lane) {
  const int mi = lane >> 3;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t b[4];
    ldsm_x4(b, sm_u32(tile + (nt2 * 16 + 8 * (mi >> 1) + (lane & 7)) * kPitch + (ks * 16 + 8 * (mi & 1)) * 2));
    mma16816(c0, a[ks], b[0], b[1]);
    mma16816(c1, a[ks], b[2], b[3]);
  }
}
// o[0..3] (four 8-wide d tiles) += A(16 x 16, from registers) * T[rows 16 kj .. +15][0..31]               (T rows = k, T cols = n)
__device__ __forceinline__ void mma_kn(float (&o)[4][4], const uint32_t (&a)[4], const uint8_t* tile, int kj, int lane) {
  const int mi = lane >> 3;
#pragma unroll
  for (int dt2 = 0; dt2 < 2; ++dt2) {
    uint32_t b[4];
    ldsm_x4_t(b, sm_u32(tile + (kj * 16 + 8 * (mi & 1) + (lane & 7)) * kPitch + (dt2 * 16 + 8 * (mi >> 1)) * 2));
    mma16816(o[2 * dt2], a, b[0], b[1]);
    mma16816(o[2 * dt2 + 1], a, b[2], b[3]);
  }
}
// 16 x 32 fp32 accumulator tile (rows row0 + g, row0 + g + 8) -> bf16 rows of `stage` (pitch kPitch), then 16-byte coalesced
// stores of the 16 rows to global (row stride `ld` elements).  Warp-private staging rows.
__device__ __forceinline__ void store_rows(const float (&o)[4][4], float s0, float s1, uint8_t* stage, int row0, __nv_bfloat16* gdst, int ld, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    *reinterpret_cast<uint32_t*>(stage + (row0 + g) * kPitch + (nt * 8 + 2 * t) * 2) = pack_bf16(o[nt][0] * s0, o[nt][1] * s0);
    *reinterpret_cast<uint32_t*>(stage + (row0 + g + 8) * kPitch + (nt * 8 + 2 * t) * 2) = pack_bf16(o[nt][2] * s1, o[nt][3] * s1);
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = lane + 32 * k, row = idx >> 2, c = idx & 3;
    *reinterpret_cast<int4*>(gdst + (size_t)(row0 + row) * ld + c * 8) = *reinterpret_cast<const int4*>(stage + (row0 + row) * kPitch + c * 16);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------ forward
template <int NP>
__global__ void __launch_bounds__(NP * 2, 2) attn_mma_fwd_kernel(AttnMmaParams p) {
  constexpr int NT = NP / 8;
  constexpr int CHF = 6;                                 // n-tiles (8 keys) per softmax chunk: 48 keys
  static_assert(NT % CHF == 0, "NP must be a multiple of 48");
  constexpr int kTile = NP * kPitch;
  extern __shared__ __align__(16) uint8_t sm[];          // 2 stages x {Q, K, V}
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.x % p.nH, cta = blockIdx.x / p.nH;
  const int ld3 = 3 * p.C;
  const float sc2 = p.scale * kL2e;
  const int i0 = warp * 16;
  auto issue = [&](int win, int st) {
    uint8_t* s = sm + st * 3 * kTile;
    const __nv_bfloat16* base = p.qkv + (size_t)win * NP * ld3 + h * MHD;
    load_tile<NP>(s, base, ld3);
    load_tile<NP>(s + kTile, base + p.C, ld3);
    load_tile<NP>(s + 2 * kTile, base + 2 * p.C, ld3);
    cp_async_commit();
  };
  if (cta < p.B_) issue(cta, 0);
  int it = 0;
  for (int win = cta; win < p.B_; win += p.per_head, ++it) {
    const int st = it & 1;
    const bool more = win + p.per_head < p.B_;
    if (more) { issue(win + p.per_head, st ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    uint8_t* sQ = sm + st * 3 * kTile;
    const uint8_t* sK = sQ + kTile;
    const uint8_t* sV = sQ + 2 * kTile;
    uint32_t aq[2][4];
    load_a_frags(aq, sQ, i0, lane);
    // online softmax over chunks of 48 keys (FlashAttention-2): 24 logit registers at a time instead of 72, so two CTAs fit an SM
    const int r0 = i0 + g, r1 = r0 + 8;
    const float* b0 = p.bias + ((size_t)h * NP + r0) * NP + 2 * t;
    const float* m0 = nullptr;
    if (p.mask != nullptr) {
      const int mw = win % p.nW;
      if (p.mask_nz == nullptr || p.mask_nz[mw]) m0 = p.mask + ((size_t)mw * NP + r0) * NP + 2 * t;
    }
    float mx0 = -INFINITY, mx1 = -INFINITY, sum0 = 0.f, sum1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll
    for (int c = 0; c < NT / CHF; ++c) {
      float s[CHF][4];
#pragma unroll
      for (int k = 0; k < CHF; ++k) { s[k][0] = s[k][1] = s[k][2] = s[k][3] = 0.f; }
#pragma unroll
      for (int k2 = 0; k2 < CHF / 2; ++k2) mma_nk(s[2 * k2], s[2 * k2 + 1], aq, sK, c * (CHF / 2) + k2, lane);
      // logits in the log2 domain: scale * s + bias (+ mask); rows r0 and r1 = r0 + 8, columns 8 nt + 2 t + {0, 1}
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int k = 0; k < CHF; ++k) {
        const int nt = c * CHF + k;
        float2 ba = __ldg(reinterpret_cast<const float2*>(b0 + nt * 8)), bb = __ldg(reinterpret_cast<const float2*>(b0 + 8 * NP + nt * 8));
        if (m0 != nullptr) {
          const float2 ma = __ldg(reinterpret_cast<const float2*>(m0 + nt * 8)), mb = __ldg(reinterpret_cast<const float2*>(m0 + 8 * NP + nt * 8));
          ba.x += ma.x; ba.y += ma.y; bb.x += mb.x; bb.y += mb.y;
        }
        s[k][0] = fmaf(s[k][0], sc2, ba.x * kL2e); s[k][1] = fmaf(s[k][1], sc2, ba.y * kL2e);
        s[k][2] = fmaf(s[k][2], sc2, bb.x * kL2e); s[k][3] = fmaf(s[k][3], sc2, bb.y * kL2e);
        cm0 = fmaxf(cm0, fmaxf(s[k][0], s[k][1])); cm1 = fmaxf(cm1, fmaxf(s[k][2], s[k][3]));
      }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float n0 = fmaxf(mx0, cm0), n1 = fmaxf(mx1, cm1);
      const float f0 = ex2a(mx0 - n0), f1 = ex2a(mx1 - n1);          // rescale of what has been accumulated so far (0 on the first chunk)
      mx0 = n0; mx1 = n1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int k = 0; k < CHF; ++k) {
        s[k][0] = ex2a(s[k][0] - n0); s[k][1] = ex2a(s[k][1] - n0);
        s[k][2] = ex2a(s[k][2] - n1); s[k][3] = ex2a(s[k][3] - n1);
        ps0 += s[k][0] + s[k][1]; ps1 += s[k][2] + s[k][3];
      }
      sum0 = fmaf(sum0, f0, ps0); sum1 = fmaf(sum1, f1, ps1);           // per-thread partial sums; the quad adds up after the loop
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { o[nt][0] *= f0; o[nt][1] *= f0; o[nt][2] *= f1; o[nt][3] *= f1; }
#pragma unroll
      for (int k2 = 0; k2 < CHF / 2; ++k2) {
        uint32_t a[4] = {pack_bf16(s[2 * k2][0], s[2 * k2][1]), pack_bf16(s[2 * k2][2], s[2 * k2][3]),
                         pack_bf16(s[2 * k2 + 1][0], s[2 * k2 + 1][1]), pack_bf16(s[2 * k2 + 1][2], s[2 * k2 + 1][3])};
        mma_kn(o, a, sV, c * (CHF / 2) + k2, lane);
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    if (t == 0) {
      float* l = p.lse + ((size_t)win * p.nH + h) * NP;
      l[r0] = (mx0 + log2f(sum0)) * 0.6931471805599453f;
      l[r1] = (mx1 + log2f(sum1)) * 0.6931471805599453f;
    }
    // this warp's Q rows are dead (only this warp read them): stage O there
    store_rows(o, 1.0f / sum0, 1.0f / sum1, sQ, i0, p.out + (size_t)win * NP * p.C + h * MHD, p.C, lane);
    __syncthreads();                                 // every warp is done with this stage before it is refilled
  }
}

// ------------------------------------------------------------------------------------------ backward
template <int NP>
__global__ void __launch_bounds__(NP * 2, 1) attn_mma_bwd_kernel(AttnMmaParams p) {
  constexpr int NT = NP / 8, NK = NP / 16;
  constexpr int CH = 6;                                  // n-tiles (8 keys) per chunk of phase 1: 48 keys
  static_assert(NT % CH == 0, "NP must be a multiple of 48");
  constexpr int kTile = NP * kPitch;
  constexpr int kPP = NP * 2 + 16;                       // pitch of the P / dS tiles (bf16 [NP][NP] + 16 B pad)
  extern __shared__ __align__(16) uint8_t sm[];          // 2 stages x {Q, K, V, dO, O}, then P, dS
  uint8_t* sP = sm + 2 * 5 * kTile;
  uint8_t* sdS = sP + NP * kPP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.x % p.nH, cta = blockIdx.x / p.nH;
  const int ld3 = 3 * p.C;
  const float sc2 = p.scale * kL2e;
  const int i0 = warp * 16;
  auto issue = [&](int win, int st) {
    uint8_t* s = sm + st * 5 * kTile;
    const __nv_bfloat16* base = p.qkv + (size_t)win * NP * ld3 + h * MHD;
    load_tile<NP>(s, base, ld3);
    load_tile<NP>(s + kTile, base + p.C, ld3);
    load_tile<NP>(s + 2 * kTile, base + 2 * p.C, ld3);
    load_tile<NP>(s + 3 * kTile, p.dout + (size_t)win * NP * p.C + h * MHD, p.C);
    load_tile<NP>(s + 4 * kTile, p.out + (size_t)win * NP * p.C + h * MHD, p.C);
    cp_async_commit();
  };
  float db[NT][4];                                       // dBias of this thread's (row, column) positions, summed over the CTA's items
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { db[nt][0] = db[nt][1] = db[nt][2] = db[nt][3] = 0.f; }
  if (cta < p.B_) issue(cta, 0);
  // the saved LSE of this thread's two rows comes from HBM: fetched one item ahead (raw -- it is scaled where it is used, so no
  // instruction waits for the load next to it)
  const int lr0 = warp * 16 + g, lr1 = lr0 + 8;
  float lse0_n = 0.f, lse1_n = 0.f;
  if (cta < p.B_) { const float* l = p.lse + ((size_t)cta * p.nH + h) * NP; lse0_n = l[lr0]; lse1_n = l[lr1]; }
  int it = 0;
  for (int win = cta; win < p.B_; win += p.per_head, ++it) {
    const int st = it & 1;
    const bool more = win + p.per_head < p.B_;
    const float lse0 = lse0_n, lse1 = lse1_n;
    if (more) {
      const float* l = p.lse + ((size_t)(win + p.per_head) * p.nH + h) * NP;
      lse0_n = l[lr0]; lse1_n = l[lr1];
    }
    if (more) { issue(win + p.per_head, st ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    uint8_t* sQ = sm + st * 5 * kTile;
    uint8_t* sK = sQ + kTile;
    uint8_t* sV = sQ + 2 * kTile;
    const uint8_t* sdO = sQ + 3 * kTile;
    uint8_t* sO = sQ + 4 * kTile;
    // ---------------- phase 1: warp = query rows i0 .. i0 + 15
    const int r0 = i0 + g, r1 = r0 + 8;
    uint32_t aq[2][4], ado[2][4];
    load_a_frags(aq, sQ, i0, lane);
    load_a_frags(ado, sdO, i0, lane);
    // D = rowsum(dO . O): thread (g, t) sums columns 8 t .. 8 t + 7 of rows r0 and r1, then the quad adds up
    float d0 = 0.f, d1 = 0.f;
    {
      const int4 x0 = *reinterpret_cast<const int4*>(sdO + r0 * kPitch + t * 16), y0 = *reinterpret_cast<const int4*>(sO + r0 * kPitch + t * 16);
      const int4 x1 = *reinterpret_cast<const int4*>(sdO + r1 * kPitch + t * 16), y1 = *reinterpret_cast<const int4*>(sO + r1 * kPitch + t * 16);
      const uint32_t* a0 = reinterpret_cast<const uint32_t*>(&x0); const uint32_t* c0 = reinterpret_cast<const uint32_t*>(&y0);
      const uint32_t* a1 = reinterpret_cast<const uint32_t*>(&x1); const uint32_t* c1 = reinterpret_cast<const uint32_t*>(&y1);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        d0 = fmaf(bf16_lo(a0[e]), bf16_lo(c0[e]), d0); d0 = fmaf(bf16_hi(a0[e]), bf16_hi(c0[e]), d0);
        d1 = fmaf(bf16_lo(a1[e]), bf16_lo(c1[e]), d1); d1 = fmaf(bf16_hi(a1[e]), bf16_hi(c1[e]), d1);
      }
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    }
    const float l0 = lse0 * kL2e, l1 = lse1 * kL2e;
    const float* b0 = p.bias + ((size_t)h * NP + r0) * NP + 2 * t;
    const float* m0 = nullptr;
    if (p.mask != nullptr) {
      const int mw = win % p.nW;
      if (p.mask_nz == nullptr || p.mask_nz[mw]) m0 = p.mask + ((size_t)mw * NP + r0) * NP + 2 * t;
    }
    float dq[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f; }
#pragma unroll
    for (int c = 0; c < NT / CH; ++c) {
      float s[CH][4], dp[CH][4];
#pragma unroll
      for (int k = 0; k < CH; ++k) { s[k][0] = s[k][1] = s[k][2] = s[k][3] = 0.f; dp[k][0] = dp[k][1] = dp[k][2] = dp[k][3] = 0.f; }
#pragma unroll
      for (int k2 = 0; k2 < CH / 2; ++k2) {
        mma_nk(s[2 * k2], s[2 * k2 + 1], aq, sK, c * (CH / 2) + k2, lane);       // S = Q K^T
        mma_nk(dp[2 * k2], dp[2 * k2 + 1], ado, sV, c * (CH / 2) + k2, lane);    // dP = dO V^T
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int nt = c * CH + k;
        float2 ba = __ldg(reinterpret_cast<const float2*>(b0 + nt * 8)), bb = __ldg(reinterpret_cast<const float2*>(b0 + 8 * NP + nt * 8));
        if (m0 != nullptr) {
          const float2 ma = __ldg(reinterpret_cast<const float2*>(m0 + nt * 8)), mb = __ldg(reinterpret_cast<const float2*>(m0 + 8 * NP + nt * 8));
          ba.x += ma.x; ba.y += ma.y; bb.x += mb.x; bb.y += mb.y;
        }
        const float p0 = ex2a(fmaf(s[k][0], sc2, ba.x * kL2e) - l0), p1 = ex2a(fmaf(s[k][1], sc2, ba.y * kL2e) - l0);
        const float p2 = ex2a(fmaf(s[k][2], sc2, bb.x * kL2e) - l1), p3 = ex2a(fmaf(s[k][3], sc2, bb.y * kL2e) - l1);
        const float e0 = p0 * (dp[k][0] - d0), e1 = p1 * (dp[k][1] - d0), e2 = p2 * (dp[k][2] - d1), e3 = p3 * (dp[k][3] - d1);
        db[nt][0] += e0; db[nt][1] += e1; db[nt][2] += e2; db[nt][3] += e3;
        // P and scale * dS as bf16: in registers as the A operand of dQ, in shared memory (rows = query i) for phase 2
        s[k][0] = p0; s[k][1] = p1; s[k][2] = p2; s[k][3] = p3;
        dp[k][0] = e0 * p.scale; dp[k][1] = e1 * p.scale; dp[k][2] = e2 * p.scale; dp[k][3] = e3 * p.scale;
        const int col = (nt * 8 + 2 * t) * 2;
        *reinterpret_cast<uint32_t*>(sP + r0 * kPP + col) = pack_bf16(p0, p1);
        *reinterpret_cast<uint32_t*>(sP + r1 * kPP + col) = pack_bf16(p2, p3);
        *reinterpret_cast<uint32_t*>(sdS + r0 * kPP + col) = pack_bf16(dp[k][0], dp[k][1]);
        *reinterpret_cast<uint32_t*>(sdS + r1 * kPP + col) = pack_bf16(dp[k][2], dp[k][3]);
      }
#pragma unroll
      for (int k2 = 0; k2 < CH / 2; ++k2) {                                      // dQ += (scale dS) K
        uint32_t a[4] = {pack_bf16(dp[2 * k2][0], dp[2 * k2][1]), pack_bf16(dp[2 * k2][2], dp[2 * k2][3]),
                         pack_bf16(dp[2 * k2 + 1][0], dp[2 * k2 + 1][1]), pack_bf16(dp[2 * k2 + 1][2], dp[2 * k2 + 1][3])};
        mma_kn(dq, a, sK, c * (CH / 2) + k2, lane);
      }
    }
    // this warp's O rows are dead (only this warp read them, for D): stage dQ there
    store_rows(dq, 1.0f, 1.0f, sO, i0, p.dqkv + (size_t)win * NP * ld3 + h * MHD, ld3, lane);
    __syncthreads();                                 // P, dS complete; every warp is done with K and V
    // ---------------- phase 2: warp = keys j0 .. j0 + 15:  dV = P^T dO,  dK = (scale dS)^T Q
    {
      const int j0 = i0, mi = lane >> 3;
      float dv[4][4], dk[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f; dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f; }
#pragma unroll
      for (int ki = 0; ki < NK; ++ki) {
        uint32_t ap[4], as[4];
        const int off = (ki * 16 + 8 * (mi >> 1) + (lane & 7)) * kPP + (j0 + 8 * (mi & 1)) * 2;
        ldsm_x4_t(ap, sm_u32(sP + off));
        ldsm_x4_t(as, sm_u32(sdS + off));
        mma_kn(dv, ap, sdO, ki, lane);
        mma_kn(dk, as, sQ, ki, lane);
      }
      // K and V are read in phase 1 only: their rows j0 .. j0 + 15 are this warp's staging for dK / dV
      store_rows(dk, 1.0f, 1.0f, sK, j0, p.dqkv + (size_t)win * NP * ld3 + p.C + h * MHD, ld3, lane);
      store_rows(dv, 1.0f, 1.0f, sV, j0, p.dqkv + (size_t)win * NP * ld3 + 2 * p.C + h * MHD, ld3, lane);
    }
    __syncthreads();                                 // this stage may be refilled (by the issue at the top of the next iteration but one)
  }
  if (cta < p.B_) {
    float* dst = p.dbias + ((size_t)h * NP + i0 + g) * NP + 2 * t;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      atomicAdd(dst + nt * 8, db[nt][0]); atomicAdd(dst + nt * 8 + 1, db[nt][1]);
      atomicAdd(dst + 8 * NP + nt * 8, db[nt][2]); atomicAdd(dst + 8 * NP + nt * 8 + 1, db[nt][3]);
    }
  }
}

int mma_common(const swin_attn_args* a, AttnMmaParams* out, bool bwd, int ctas_per_sm) {
  SWIN_REQUIRE(a->ws == 12, "attn(mma): window 12 only (got %d)", a->ws);
  SWIN_REQUIRE(a->B_ >= 0 && a->nH > 0, "attn: bad shape");
  SWIN_REQUIRE(a->qkv && a->bias && a->lse && a->out, "attn: null pointer");
  SWIN_REQUIRE(a->mask == nullptr || (a->nW > 0 && a->B_ % a->nW == 0), "attn: B_ must be a multiple of nW when a mask is given");
  SWIN_REQUIRE(aligned16(a->qkv) && aligned16(a->out), "attn: alignment");
  if (bwd) SWIN_REQUIRE(a->dout && a->dqkv && a->dbias && aligned16(a->dout) && aligned16(a->dqkv), "attn_bwd: null/misaligned pointer");
  AttnMmaParams p;
  p.B_ = a->B_; p.nH = a->nH; p.nW = a->nW > 0 ? a->nW : 1; p.C = a->nH * MHD; p.scale = a->scale;
  int per_head = (persistent_sms() * ctas_per_sm) / a->nH;
  if (per_head > a->B_) per_head = a->B_;
  if (per_head < 1) per_head = 1;
  p.per_head = per_head;
  p.qkv = (const __nv_bfloat16*)a->qkv; p.bias = a->bias; p.mask = a->mask; p.mask_nz = a->mask ? a->mask_nz : nullptr;
  p.out = (__nv_bfloat16*)a->out; p.lse = a->lse;
  p.dout = (const __nv_bfloat16*)a->dout; p.dqkv = (__nv_bfloat16*)a->dqkv; p.dbias = a->dbias;
  *out = p;
  return 0;
}

}  // namespace

int attn_mma_supported(int ws) { return ws == 12; }

int attn_mma_fwd(const swin_attn_args* a, cudaStream_t st) {
  AttnMmaParams p;
  int rc = mma_common(a, &p, false, 2);      // two CTAs per SM
  if (rc) return rc;
  if (p.B_ == 0) return 0;
  constexpr int NP = 144;
  const size_t smem = 2 * 3 * NP * kPitch;
  rc = ensure_dyn_smem((const void*)attn_mma_fwd_kernel<NP>, (int)smem);
  if (rc) return rc;
  attn_mma_fwd_kernel<NP><<<p.nH * p.per_head, NP * 2, smem, st>>>(p);
  SWIN_LAUNCH_CHECK();
  return 0;
}

int attn_mma_bwd(const swin_attn_args* a, cudaStream_t st) {
  AttnMmaParams p;
  int rc = mma_common(a, &p, true, 1);
  if (rc) return rc;
  if (p.B_ == 0) return 0;
  constexpr int NP = 144;
  const size_t smem = 2 * 5 * NP * kPitch + 2 * NP * (NP * 2 + 16);
  rc = ensure_dyn_smem((const void*)attn_mma_bwd_kernel<NP>, (int)smem);
  if (rc) return rc;
  attn_mma_bwd_kernel<NP><<<p.nH * p.per_head, NP * 2, smem, st>>>(p);
  SWIN_LAUNCH_CHECK();
  return 0;
}

}  // namespace swin
