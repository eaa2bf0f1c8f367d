// HBM-bound kernels of the Swin path: bit-exact index ops (shift / partition / reverse / mask),
// the LayerNorm family (plain, fused with pad+roll+partition, fused with the PatchMerging gather),
// casts, column sums and the relative-position-bias expand / reduce.
// All accesses are 128-bit and coalesced per row; index arithmetic is done once per row.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace swin {

// ------------------------------------------------------------------------------------------
// Row copy machinery: a group of L lanes (L = power of two <= 32 dividing the row's 16-byte
// vector count) moves one row; a warp therefore moves 32/L rows at a time.
// ------------------------------------------------------------------------------------------
static inline int lanes_per_row(int vecs_per_row) {
  int l = 1;
  while (l < 32 && (vecs_per_row % (l * 2)) == 0) l *= 2;
  return l;
}

enum { MAP_PARTITION = 0, MAP_REVERSE = 1, MAP_GATHER = 2, MAP_SCATTER = 3 };

// One launch covers `rows` destination rows; src row index (or -1 => zeros) comes from the map.
template <int MAP>
__global__ void __launch_bounds__(256) row_map_copy_kernel(const int4* __restrict__ src, int4* __restrict__ dst,
                                                           WinGeom g, int vpr, int L, int rows) {
  // rows < 2^30 and rows * vpr < 2^31 (checked by the launcher): all index math is 32-bit with multiply-high divisions
  const int groups_per_block = blockDim.x / L;
  const int gl = threadIdx.x % L;
  int row = blockIdx.x * groups_per_block + threadIdx.x / L;
  const int stride = gridDim.x * groups_per_block;
  const int per_img_slots = g.nW * g.N;
  const int per_img_tok = g.H * g.W;
  for (; row < rows; row += stride) {
    int srow, in;
    if (MAP == MAP_GATHER) {            // dst = window slots, src = tokens
      int b = split_slot_row(g, row, &in);
      int t = slot_to_token(g, in);
      srow = t < 0 ? -1 : b * per_img_tok + t;
    } else if (MAP == MAP_SCATTER) {    // dst = tokens, src = window slots (always valid)
      int b = split_tok_row(g, row, &in);
      srow = b * per_img_slots + token_to_slot(g, in);
    } else if (MAP == MAP_PARTITION) {  // H,W already padded (H==Hp), shift 0
      int b = split_slot_row(g, row, &in);
      srow = b * per_img_tok + slot_to_token(g, in);
    } else {                            // MAP_REVERSE
      int b = split_tok_row(g, row, &in);
      srow = b * per_img_slots + token_to_slot(g, in);
    }
    int4* d = dst + (long long)row * vpr;
    if (srow < 0) {
      for (int v = gl; v < vpr; v += L) d[v] = make_int4(0, 0, 0, 0);
    } else {
      const int4* s = src + (long long)srow * vpr;
      for (int v = gl; v < vpr; v += L) d[v] = __ldg(s + v);
    }
  }
}

template <int MAP>
static int launch_row_map(const void* src, void* dst, const WinGeom& g, int elem_bytes, long long rows, cudaStream_t st) {
  SWIN_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "elem_bytes must be 2 or 4");
  SWIN_REQUIRE(((long long)g.C * elem_bytes) % 16 == 0, "row bytes (C*elem) must be a multiple of 16");
  SWIN_REQUIRE(aligned16(src) && aligned16(dst), "pointers must be 16-byte aligned");
  SWIN_REQUIRE(g.B > 0 && g.H > 0 && g.W > 0 && g.ws > 0 && g.shift >= 0 && g.shift < g.ws, "bad geometry");
  if (rows == 0) return 0;
  SWIN_REQUIRE(rows < (1ll << 30) && rows * (g.C * elem_bytes / 16) < (1ll << 31), "tensor too large for the gather kernels' 32-bit indices");
  int vpr = g.C * elem_bytes / 16;
  int L = lanes_per_row(vpr);
  int gpb = 256 / L;
  long long blocks = (rows + gpb - 1) / gpb;
  int grid = (int)(blocks < (long long)kNumSMs * 16 ? blocks : (long long)kNumSMs * 16);
  row_map_copy_kernel<MAP><<<grid, 256, 0, st>>>((const int4*)src, (int4*)dst, g, vpr, L, (int)rows);
  SWIN_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// shift mask (REF:370-389): region ids on the padded grid, in window-frame coordinates.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int region1d(int p, int P, int ws, int shift) { return (p >= P - ws) + (p >= P - shift); }

__global__ void shift_mask_kernel(float* __restrict__ mask, WinGeom g) {
  long long total = (long long)g.nW * g.N * g.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int q = (int)(i % g.N);
    int p = (int)((i / g.N) % g.N);
    int w = (int)(i / ((long long)g.N * g.N));
    int wh = w / g.nww, ww = w - wh * g.nww;
    int rp = 3 * region1d(wh * g.ws + p / g.ws, g.Hp, g.ws, g.shift) + region1d(ww * g.ws + p % g.ws, g.Wp, g.ws, g.shift);
    int rq = 3 * region1d(wh * g.ws + q / g.ws, g.Hp, g.ws, g.shift) + region1d(ww * g.ws + q % g.ws, g.Wp, g.ws, g.shift);
    mask[i] = (rp == rq) ? 0.0f : -100.0f;
  }
}

__global__ void mask_nonzero_kernel(const float* __restrict__ mask, int* __restrict__ flags, int nW, int NN) {
  const int w = blockIdx.x;
  int any = 0;
  for (int e = threadIdx.x; e < NN; e += blockDim.x) any |= (mask[(size_t)w * NN + e] != 0.0f);
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flags[w] = any ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// relative position bias: table ((2ws-1)^2, nH) <-> dense (nH, N, N)   (REF:101-111, :135-137)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int rel_index(int i, int j, int ws) {
  return (i / ws - j / ws + ws - 1) * (2 * ws - 1) + (i % ws - j % ws + ws - 1);
}
__global__ void rel_bias_expand_kernel(const float* __restrict__ table, float* __restrict__ bias, int nH, int ws) {
  int N = ws * ws;
  int total = nH * N * N;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    int j = e % N, i = (e / N) % N, h = e / (N * N);
    bias[e] = table[rel_index(i, j, ws) * nH + h];
  }
}
// one WARP per table entry: deterministic gather-sum of every (i,j) that maps to it -- the (rj, cj) pairs of an entry are spread
// over the lanes (fixed assignment, fixed shuffle tree), so the result does not depend on scheduling
__global__ void rel_bias_reduce_kernel(const float* __restrict__ dbias, float* __restrict__ dtable, int nH, int ws) {
  const int R = 2 * ws - 1;
  const int total = R * R * nH;
  const int N = ws * ws;
  const int lane = threadIdx.x & 31;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total; e += (gridDim.x * blockDim.x) >> 5) {
    const int h = e % nH, r = e / nH;
    const int dr = r / R - (ws - 1), dc = r % R - (ws - 1);   // ri - rj, ci - cj
    float s = 0.f;
    for (int t = lane; t < N; t += 32) {
      const int rj = t / ws, cj = t - rj * ws;
      const int ri = rj + dr, ci = cj + dc;
      if (ri >= 0 && ri < ws && ci >= 0 && ci < ws) s += dbias[((size_t)h * N + (ri * ws + ci)) * N + t];
    }
    s = warp_sum(s);
    if (lane == 0) dtable[e] += s;
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm family.  One warp per LN row; the row is NSEG gathered segments of C floats
// (NSEG = 1 for plain / window mode, 4 for PatchMerging).  VPL = float4 vectors held per lane.
// ------------------------------------------------------------------------------------------
struct LnGeom {
  WinGeom g;       // window mode
  int mode;        // 0 plain, 1 window, 2 merge
  int C;           // segment width
  int nseg;        // 1 or 4
  int H2, W2;      // merge
  FastDiv dper2, dW2, dvps;   // dividers: H2*W2, W2, C/4
  int rows;        // iteration rows (see kernels); < 2^30, and rows * row width < 2^31 float4 (checked in ln_geom)
  float eps;
  // backward, modes 0 and 1 (iteration rows = tokens): optional second output  y2[slot(token)] = y2_scale[b] * dx[token]  (window-slot layout,
  // dtype of dy) + its column sums: the dY of the proj Linear, produced while dx is still in registers
  WinGeom g2;
  void* y2;
  const float* y2_scale;
  float* y2_colsum;
};

// merged row (b, oh, ow) segment q -> source token row or -1   (REF:288-292 order (0,0),(1,0),(0,1),(1,1))
__device__ __forceinline__ int merge_src(const LnGeom& lg, int row, int q) {
  int per = lg.H2 * lg.W2;
  int b = fdiv(row, lg.dper2), r = row - b * per;
  int oh = fdiv(r, lg.dW2), ow = r - oh * lg.W2;
  int h = 2 * oh + (q & 1), w = 2 * ow + (q >> 1);
  return (h < lg.g.H && w < lg.g.W) ? (b * lg.g.H + h) * lg.g.W + w : -1;
}

template <typename T> struct Vec4IO;
template <> struct Vec4IO<float> {
  typedef float4 raw_t;
  static __device__ __forceinline__ raw_t ldraw(const float* p, long long v) { return __ldg(reinterpret_cast<const float4*>(p) + v); }
  static __device__ __forceinline__ float4 cvt(raw_t r) { return r; }
  static __device__ __forceinline__ float4 ld(const float* p, long long v) { return __ldg(reinterpret_cast<const float4*>(p) + v); }
  static __device__ __forceinline__ void st(float* p, long long v, float4 x) { reinterpret_cast<float4*>(p)[v] = x; }
};
template <> struct Vec4IO<__nv_bfloat16> {
  typedef uint2 raw_t;
  static __device__ __forceinline__ raw_t ldraw(const __nv_bfloat16* p, long long v) { return __ldg(reinterpret_cast<const uint2*>(p) + v); }
  static __device__ __forceinline__ float4 cvt(raw_t u) { return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y)); }
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p, long long v) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(p) + v);
    return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, long long v, float4 x) {
    reinterpret_cast<uint2*>(p)[v] = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
  }
};

// A row of the LN is handled by a group of G lanes (G = 8, 16 or 32) holding VPL float4 each, so a warp works on
// 32/G rows at once: for the narrow stage-0/1 rows (C = 96, 192) this quadruples / doubles the loads in flight
// per warp, which is what these HBM-bound kernels are limited by.
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Forward.  Iteration rows: mode 0 -> LN rows; mode 1 -> window SLOTS (pad slots get zeros);
// mode 2 -> merged rows.
template <int VPL, int G, typename YT>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, YT* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, LnGeom lg) {
  constexpr int R = 32 / G;                  // rows per warp
  const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
  const int wpb = blockDim.x >> 5;
  const int vps = lg.C >> 2;                 // float4 per segment
  const int vrow = vps * lg.nseg;            // float4 per LN row
  const float inv_n = 1.0f / (float)(lg.C * lg.nseg);
  const int per_img_tok = lg.g.H * lg.g.W;
  for (int base = (blockIdx.x * wpb + (threadIdx.x >> 5)) * R; base < lg.rows; base += gridDim.x * wpb * R) {
    const int row = base + gi;
    const bool inr = row < lg.rows;
    int stat_row = row, src0 = row;
    bool pad = false;
    if (lg.mode == 1 && inr) {
      int in;
      int b = split_slot_row(lg.g, row, &in);
      int t = slot_to_token(lg.g, in);
      pad = t < 0;                           // zero padding AFTER the norm (REF:211 then :218)
      src0 = b * per_img_tok + (pad ? 0 : t);
      stat_row = src0;
    }
    const bool act = inr && !pad;
    float4 r[VPL];
    bool have[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {            // all loads first (clamped index), then the sums
      const int v = gl + G * k;
      int srow = src0;
      int off = v;
      have[k] = act && v < vrow;
      if (lg.mode == 2 && have[k]) { int q = fdiv(v, lg.dvps); off = v - q * vps; srow = merge_src(lg, row, q); }
      have[k] = have[k] && srow >= 0;
      r[k] = Vec4IO<float>::ld(x, have[k] ? (long long)(srow * vps + off) : 0);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (!have[k]) r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      s += r[k].x + r[k].y + r[k].z + r[k].w;
    }
    const float mu = group_sum<G>(s) * inv_n;
    float q2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (gl + G * k < vrow) {
        float a = r[k].x - mu, b2 = r[k].y - mu, c = r[k].z - mu, d = r[k].w - mu;
        q2 += a * a + b2 * b2 + c * c + d * d;
      }
    }
    const float rs = rsqrtf(group_sum<G>(q2) * inv_n + lg.eps);
    if (act && gl == 0) { mean[stat_row] = mu; rstd[stat_row] = rs; }
    if (!inr) continue;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = gl + G * k;
      if (v < vrow) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!pad) {
          float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + v);
          float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + v);
          o.x = (r[k].x - mu) * rs * gm.x + bt.x;
          o.y = (r[k].y - mu) * rs * gm.y + bt.y;
          o.z = (r[k].z - mu) * rs * gm.z + bt.z;
          o.w = (r[k].w - mu) * rs * gm.w + bt.w;
        }
        Vec4IO<YT>::st(y, (long long)(row * vrow + v), o);
      }
    }
  }
}

// Second output of the backward kernels (dY of the proj Linear in window-slot layout): tokens fill their own slots, and
// the padding slots — the two rectangles hs >= H and ws >= W of the padded grid, 1-2 % of the rows — are zeroed here, so
// the caller does not have to memset the whole tensor first.
template <typename YT>
__device__ __forceinline__ void zero_pad_slots(const WinGeom& g, YT* __restrict__ y2, int vrow) {
  const int padH = g.Hp - g.H, padW = g.Wp - g.W;
  const int per_img = padH * g.Wp + g.H * padW;
  const int total = g.B * per_img;
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int pidx = blockIdx.x * wpb + (threadIdx.x >> 5); pidx < total; pidx += gridDim.x * wpb) {
    const int b = pidx / per_img;
    int q = pidx - b * per_img, hs, wsrc;
    if (q < padH * g.Wp) { hs = g.H + q / g.Wp; wsrc = q % g.Wp; }
    else { q -= padH * g.Wp; hs = q / padW; wsrc = g.W + q % padW; }
    int hh = hs - g.shift; if (hh < 0) hh += g.Hp;
    int wq = wsrc - g.shift; if (wq < 0) wq += g.Wp;
    const int wh = hh / g.ws, i = hh - wh * g.ws, ww = wq / g.ws, j = wq - ww * g.ws;
    const long long slot = (long long)b * (g.nW * g.N) + (wh * g.nww + ww) * g.N + i * g.ws + j;
    for (int v = lane; v < vrow; v += 32) Vec4IO<YT>::st(y2, slot * vrow + v, make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

// Backward.  Iteration rows: mode 0 -> LN rows; mode 1 -> TOKENS (dy read through token->slot);
// mode 2 -> merged rows (dx scattered to the 4 source tokens; pad segments dropped).
// dx = (dres) + rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma += dy*xhat; dbeta += dy.
template <int VPL, int G, typename YT>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const YT* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const float* __restrict__ dres,
                                                     float* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, LnGeom lg) {
  extern __shared__ float sred[];            // [2][vrow*4] block partials
  constexpr int R = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
  const int wpb = blockDim.x >> 5;
  const int vps = lg.C >> 2;
  const int vrow = vps * lg.nseg;
  const int width = vrow * 4;
  const float inv_n = 1.0f / (float)width;
  const int per_img_slots = lg.g.nW * lg.g.N;
  for (int i = threadIdx.x; i < 2 * width; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float4 ag[VPL], ab[VPL], a2[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) { ag[k] = make_float4(0.f, 0.f, 0.f, 0.f); ab[k] = ag[k]; a2[k] = ag[k]; }
  for (int base = (blockIdx.x * wpb + (threadIdx.x >> 5)) * R; base < lg.rows; base += gridDim.x * wpb * R) {
    const int row = base + gi;
    const bool inr = row < lg.rows;
    int dyrow = row;
    if (lg.mode == 1 && inr) {
      int in;
      int b = split_tok_row(lg.g, row, &in);
      dyrow = b * per_img_slots + token_to_slot(lg.g, in);
    }
    const float mu = inr ? mean[row] : 0.f, rs = inr ? rstd[row] : 0.f;
    // Phase A: every load of the row is issued unconditionally (clamped index, read-only path) before any use, so
    // they overlap; a load-then-use sequence per vector gets serialised by in-order issue.
    float4 xh[VPL], gd[VPL], rr[VPL];
    typename Vec4IO<YT>::raw_t draw[VPL];
    int srow_k[VPL];
    const float* rsrc = dres != nullptr ? dres : x;
    const float rflag = dres != nullptr ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = gl + G * k;
      const bool on = inr && v < vrow;
      int srow = row; int off = v;
      if (lg.mode == 2 && on) { int q = fdiv(v, lg.dvps); off = v - q * vps; srow = merge_src(lg, row, q); }
      srow_k[k] = (on && srow >= 0) ? srow * vps + off : -1;
      const long long xi = srow_k[k] >= 0 ? srow_k[k] : 0;
      xh[k] = Vec4IO<float>::ld(x, xi);
      rr[k] = Vec4IO<float>::ld(rsrc, xi);
      draw[k] = Vec4IO<YT>::ldraw(dy, on ? (long long)(dyrow * vrow + v) : 0);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = gl + G * k;
      const bool on = inr && v < vrow;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + (on ? v : 0));
      float4 d = Vec4IO<YT>::cvt(draw[k]);
      if (!on) d = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 xv = srow_k[k] >= 0 ? xh[k] : make_float4(mu, mu, mu, mu);      // pad segment: x = 0 -> handled below
      if (srow_k[k] < 0 && on) xh[k] = make_float4(-mu * rs, -mu * rs, -mu * rs, -mu * rs);   // merge-mode zero pad: xhat = (0 - mu) * rs
      else xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      if (!on) xh[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      rr[k] = make_float4(rr[k].x * rflag, rr[k].y * rflag, rr[k].z * rflag, rr[k].w * rflag);
      gd[k] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += gd[k].x + gd[k].y + gd[k].z + gd[k].w;
      s2 += gd[k].x * xh[k].x + gd[k].y * xh[k].y + gd[k].z * xh[k].z + gd[k].w * xh[k].w;
      ag[k].x += d.x * xh[k].x; ag[k].y += d.y * xh[k].y; ag[k].z += d.z * xh[k].z; ag[k].w += d.w * xh[k].w;
      ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
    }
    const float m1 = group_sum<G>(s1) * inv_n, m2 = group_sum<G>(s2) * inv_n;
    int slot2 = 0;
    float sc2 = 1.0f;
    if (lg.y2 != nullptr && inr) {
      int in2;
      const int b2 = split_tok_row(lg.g2, row, &in2);
      slot2 = b2 * (lg.g2.nW * lg.g2.N) + token_to_slot(lg.g2, in2);
      if (lg.y2_scale != nullptr) sc2 = lg.y2_scale[b2];
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (srow_k[k] >= 0) {
        float4 o;
        o.x = rs * (gd[k].x - m1 - xh[k].x * m2);
        o.y = rs * (gd[k].y - m1 - xh[k].y * m2);
        o.z = rs * (gd[k].z - m1 - xh[k].z * m2);
        o.w = rs * (gd[k].w - m1 - xh[k].w * m2);
        o.x += rr[k].x; o.y += rr[k].y; o.z += rr[k].z; o.w += rr[k].w;
        Vec4IO<float>::st(dx, srow_k[k], o);
        if (lg.y2 != nullptr) {
          o.x *= sc2; o.y *= sc2; o.z *= sc2; o.w *= sc2;
          Vec4IO<YT>::st(reinterpret_cast<YT*>(lg.y2), (long long)(slot2 * vrow + (gl + G * k)), o);
          a2[k].x += o.x; a2[k].y += o.y; a2[k].z += o.z; a2[k].w += o.w;
        }
      }
    }
  }
  if (lg.y2 != nullptr) zero_pad_slots<YT>(lg.g2, reinterpret_cast<YT*>(lg.y2), vrow);
  // block reduce of dgamma / dbeta partials, then one atomic per column per block
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v = gl + G * k;
    if (v < vrow) {
      atomicAdd(&sred[v * 4 + 0], ag[k].x); atomicAdd(&sred[v * 4 + 1], ag[k].y);
      atomicAdd(&sred[v * 4 + 2], ag[k].z); atomicAdd(&sred[v * 4 + 3], ag[k].w);
      atomicAdd(&sred[width + v * 4 + 0], ab[k].x); atomicAdd(&sred[width + v * 4 + 1], ab[k].y);
      atomicAdd(&sred[width + v * 4 + 2], ab[k].z); atomicAdd(&sred[width + v * 4 + 3], ab[k].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    atomicAdd(dgamma + i, sred[i]);
    atomicAdd(dbeta + i, sred[width + i]);
  }
  if (lg.y2_colsum != nullptr) {               // column sums of the second output, reusing the first half of sred
    __syncthreads();
    for (int i = threadIdx.x; i < width; i += blockDim.x) sred[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = gl + G * k;
      if (v < vrow) {
        atomicAdd(&sred[v * 4 + 0], a2[k].x); atomicAdd(&sred[v * 4 + 1], a2[k].y);
        atomicAdd(&sred[v * 4 + 2], a2[k].z); atomicAdd(&sred[v * 4 + 3], a2[k].w);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < width; i += blockDim.x) atomicAdd(lg.y2_colsum + i, sred[i]);
  }
}

// Prefetching variant of the backward kernel for narrow rows (<= 4 float4 per lane): the row loads (x, residual gradient,
// dy) of iteration i+1 are issued with cp.async into a per-warp, lane-private smem ring BEFORE iteration i is processed,
// so two iterations' worth of bytes are in flight per warp without holding registers for them (the register-held
// version is capped at 16 warps x 3.8 KB per SM by its 126 registers).  A lane only reads back slots it filled itself.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int VPL, typename YT> struct LnBwdStage {
  static constexpr int kRaw = sizeof(typename Vec4IO<YT>::raw_t);            // 16 (fp32 dy) or 8 (bf16 dy)
  static constexpr int kBytes = VPL * 32 * (16 + 16 + kRaw);                 // one stage of one warp
};

template <int VPL, int G, typename YT>
__global__ void __launch_bounds__(256) ln_bwd_pf_kernel(const YT* __restrict__ dy, const float* __restrict__ x,
                                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const float* __restrict__ dres,
                                                        float* __restrict__ dx, float* __restrict__ dgamma,
                                                        float* __restrict__ dbeta, LnGeom lg) {
  extern __shared__ __align__(16) float sred[];            // [2][vrow*4] block partials, then the staging ring
  typedef LnBwdStage<VPL, YT> Stg;
  typedef typename Vec4IO<YT>::raw_t raw_t;
  constexpr int R = 32 / G;
  const int lane = threadIdx.x & 31, gl = lane % G, gi = lane / G;
  const int wpb = blockDim.x >> 5;
  const int vps = lg.C >> 2;
  const int vrow = vps * lg.nseg;
  const int width = vrow * 4;
  const float inv_n = 1.0f / (float)width;
  const int per_img_slots = lg.g.nW * lg.g.N;
  uint8_t* stg = reinterpret_cast<uint8_t*>(sred) + (((size_t)2 * width * sizeof(float) + 15) & ~(size_t)15) +
                 (size_t)(threadIdx.x >> 5) * 2 * Stg::kBytes;
  for (int i = threadIdx.x; i < 2 * width; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float4 ag[VPL], ab[VPL], a2[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) { ag[k] = make_float4(0.f, 0.f, 0.f, 0.f); ab[k] = ag[k]; a2[k] = ag[k]; }
  const float* rsrc = dres != nullptr ? dres : x;
  const bool has_res = dres != nullptr;
  const int step = gridDim.x * wpb * R;
  const int base0 = (blockIdx.x * wpb + (threadIdx.x >> 5)) * R;

  // float4 offsets of this lane's VPL vectors of row `row` (-1: nothing to read / write) and its dy row
  auto locate = [&](int row, bool inr, int* srow_k, int* dyrow) {
    *dyrow = row;
    if (lg.mode == 1 && inr) {
      int in;
      int b = split_tok_row(lg.g, row, &in);
      *dyrow = b * per_img_slots + token_to_slot(lg.g, in);
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = gl + G * k;
      const bool on = inr && v < vrow;
      int srow = row, off = v;
      if (lg.mode == 2 && on) { int q = fdiv(v, lg.dvps); off = v - q * vps; srow = merge_src(lg, row, q); }
      srow_k[k] = (on && srow >= 0) ? srow * vps + off : -1;
    }
  };
  auto issue = [&](int base, int st, float* mu_o, float* rs_o) {
    const int row = base + gi;
    const bool inr = row < lg.rows;
    int sk[VPL], dyrow;
    locate(row, inr, sk, &dyrow);
    uint8_t* sb = stg + st * Stg::kBytes;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = gl + G * k;
      const bool on = inr && v < vrow;
      // idle lanes / pad segments issue nothing (their slots are never read): a clamped address would make every such
      // lane hit one L2 line, and cp.async.cg bypasses L1 (measured on a forward variant: 40 % slower)
      if (sk[k] >= 0) {
        cp_async16(sb + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(x) + sk[k]);
        cp_async16(sb + VPL * 512 + (k * 32 + lane) * 16, reinterpret_cast<const float4*>(rsrc) + sk[k]);
      }
      if (on) {
        const raw_t* dsrc = reinterpret_cast<const raw_t*>(dy) + (long long)(dyrow * vrow + v);
        if (Stg::kRaw == 16) cp_async16(sb + VPL * 1024 + (k * 32 + lane) * 16, dsrc);
        else cp_async8(sb + VPL * 1024 + (k * 32 + lane) * 8, dsrc);
      }
    }
    *mu_o = inr ? __ldg(mean + row) : 0.f;
    *rs_o = inr ? __ldg(rstd + row) : 0.f;
  };

  float mu_n = 0.f, rs_n = 0.f;
  if (base0 < lg.rows) issue(base0, 0, &mu_n, &rs_n);
  cp_async_commit();
  int it = 0;
  for (int base = base0; base < lg.rows; base += step, ++it) {
    const float mu = mu_n, rs = rs_n;
    if (base + step < lg.rows) issue(base + step, (it + 1) & 1, &mu_n, &rs_n);
    cp_async_commit();
    cp_async_wait<1>();                         // this iteration's group has landed (the prefetch may still be in flight)
    const int row = base + gi;
    const bool inr = row < lg.rows;
    int srow_k[VPL], dyrow;
    locate(row, inr, srow_k, &dyrow);
    const uint8_t* sb = stg + (it & 1) * Stg::kBytes;
    float4 xh[VPL], gd[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = gl + G * k;
      const bool on = inr && v < vrow;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + (on ? v : 0));
      const float4 xv = *reinterpret_cast<const float4*>(sb + (k * 32 + lane) * 16);
      raw_t draw;
      if (Stg::kRaw == 16) draw = *reinterpret_cast<const raw_t*>(sb + VPL * 1024 + (k * 32 + lane) * 16);
      else draw = *reinterpret_cast<const raw_t*>(sb + VPL * 1024 + (k * 32 + lane) * 8);
      float4 d = Vec4IO<YT>::cvt(draw);
      if (!on) d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (srow_k[k] >= 0) xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      else if (on) xh[k] = make_float4(-mu * rs, -mu * rs, -mu * rs, -mu * rs);      // merge-mode zero pad: xhat = (0 - mu) * rs
      else xh[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      gd[k] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += gd[k].x + gd[k].y + gd[k].z + gd[k].w;
      s2 += gd[k].x * xh[k].x + gd[k].y * xh[k].y + gd[k].z * xh[k].z + gd[k].w * xh[k].w;
      ag[k].x += d.x * xh[k].x; ag[k].y += d.y * xh[k].y; ag[k].z += d.z * xh[k].z; ag[k].w += d.w * xh[k].w;
      ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
    }
    const float m1 = group_sum<G>(s1) * inv_n, m2 = group_sum<G>(s2) * inv_n;
    int slot2 = 0;
    float sc2 = 1.0f;
    if (lg.y2 != nullptr && inr) {
      int in2;
      const int b2 = split_tok_row(lg.g2, row, &in2);
      slot2 = b2 * (lg.g2.nW * lg.g2.N) + token_to_slot(lg.g2, in2);
      if (lg.y2_scale != nullptr) sc2 = lg.y2_scale[b2];
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (srow_k[k] >= 0) {
        float4 o;
        o.x = rs * (gd[k].x - m1 - xh[k].x * m2);
        o.y = rs * (gd[k].y - m1 - xh[k].y * m2);
        o.z = rs * (gd[k].z - m1 - xh[k].z * m2);
        o.w = rs * (gd[k].w - m1 - xh[k].w * m2);
        if (has_res) {
          const float4 rr = *reinterpret_cast<const float4*>(sb + VPL * 512 + (k * 32 + lane) * 16);
          o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        Vec4IO<float>::st(dx, (long long)srow_k[k], o);
        if (lg.y2 != nullptr) {
          o.x *= sc2; o.y *= sc2; o.z *= sc2; o.w *= sc2;
          Vec4IO<YT>::st(reinterpret_cast<YT*>(lg.y2), (long long)(slot2 * vrow + (gl + G * k)), o);
          a2[k].x += o.x; a2[k].y += o.y; a2[k].z += o.z; a2[k].w += o.w;
        }
      }
    }
  }
  cp_async_wait<0>();
  if (lg.y2 != nullptr) zero_pad_slots<YT>(lg.g2, reinterpret_cast<YT*>(lg.y2), vrow);
  // block reduce of dgamma / dbeta partials, then one atomic per column per block
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v = gl + G * k;
    if (v < vrow) {
      atomicAdd(&sred[v * 4 + 0], ag[k].x); atomicAdd(&sred[v * 4 + 1], ag[k].y);
      atomicAdd(&sred[v * 4 + 2], ag[k].z); atomicAdd(&sred[v * 4 + 3], ag[k].w);
      atomicAdd(&sred[width + v * 4 + 0], ab[k].x); atomicAdd(&sred[width + v * 4 + 1], ab[k].y);
      atomicAdd(&sred[width + v * 4 + 2], ab[k].z); atomicAdd(&sred[width + v * 4 + 3], ab[k].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    atomicAdd(dgamma + i, sred[i]);
    atomicAdd(dbeta + i, sred[width + i]);
  }
  if (lg.y2_colsum != nullptr) {               // column sums of the second output, reusing the first half of sred
    __syncthreads();
    for (int i = threadIdx.x; i < width; i += blockDim.x) sred[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = gl + G * k;
      if (v < vrow) {
        atomicAdd(&sred[v * 4 + 0], a2[k].x); atomicAdd(&sred[v * 4 + 1], a2[k].y);
        atomicAdd(&sred[v * 4 + 2], a2[k].z); atomicAdd(&sred[v * 4 + 3], a2[k].w);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < width; i += blockDim.x) atomicAdd(lg.y2_colsum + i, sred[i]);
  }
}

static int ln_geom(const swin_ln_args* a, bool bwd, LnGeom* out) {
  LnGeom lg;
  SWIN_REQUIRE(a->mode >= 0 && a->mode <= 2, "ln: bad mode %d", a->mode);
  SWIN_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->C > 0 && a->C % 4 == 0, "ln: bad shape");
  int ws = a->mode == 1 ? a->ws : 1, shift = a->mode == 1 ? a->shift : 0;
  SWIN_REQUIRE(ws > 0 && shift >= 0 && shift < ws, "ln: bad window geometry");
  lg.g = make_geom(a->B, a->H, a->W, a->C, ws, shift);
  lg.mode = a->mode; lg.C = a->C; lg.nseg = a->mode == 2 ? 4 : 1;
  lg.H2 = (a->H + 1) / 2; lg.W2 = (a->W + 1) / 2;
  lg.eps = a->eps;
  long long tokens = (long long)a->B * a->H * a->W, rows;
  if (a->mode == 0) rows = tokens;
  else if (a->mode == 1) rows = bwd ? tokens : (long long)a->B * lg.g.nW * lg.g.N;
  else rows = (long long)a->B * lg.H2 * lg.W2;
  const long long slots = (long long)a->B * lg.g.nW * lg.g.N;
  SWIN_REQUIRE(rows < (1ll << 30) && (slots > tokens ? slots : tokens) * (a->C / 4) * lg.nseg < (1ll << 31),
               "ln: tensor too large for the kernels' 32-bit row / vector indices");
  lg.rows = (int)rows;
  lg.dper2 = make_fastdiv(lg.H2 * lg.W2); lg.dW2 = make_fastdiv(lg.W2); lg.dvps = make_fastdiv(a->C / 4);
  SWIN_REQUIRE(a->y_dtype == SWIN_F32 || (a->y_dtype == SWIN_BF16), "ln: bad y dtype");
  lg.g2 = lg.g; lg.y2 = nullptr; lg.y2_scale = nullptr; lg.y2_colsum = nullptr;
  if (bwd && a->dy2 != nullptr) {
    SWIN_REQUIRE(a->mode == 0 || a->mode == 1, "ln_bwd: the second output (dy2) is available in modes 0 and 1 (rows = tokens)");
    SWIN_REQUIRE(a->ws2 > 0 && a->shift2 >= 0 && a->shift2 < a->ws2 && aligned16(a->dy2), "ln_bwd: bad dy2 geometry/alignment");
    lg.g2 = make_geom(a->B, a->H, a->W, a->C, a->ws2, a->shift2);
    lg.y2 = a->dy2; lg.y2_scale = a->dy2_scale; lg.y2_colsum = a->dy2_colsum;
  }
  *out = lg;
  return 0;
}

// lanes per row: the smallest of 8/16/32 that keeps <= 4 float4 per lane (else 32 with more per lane)

// LayerNorm backward for WIDE merged rows (PatchMerging, mode 2, 4C = 768 .. 2048): the whole 256-thread block owns one row at a
// time (thread = one or two float4 columns), so the per-thread state is a handful of registers instead of the 250+ (with
// spills) the lane-group kernel needs at 12-16 float4 per lane -- that one launch ran at 0.6-2.1 TB/s.  Row statistics go
// through a block reduction; dgamma / dbeta partials stay in registers (a thread keeps its columns) until the end.
template <typename YT>
__global__ void __launch_bounds__(256) ln_bwd_block_kernel(const YT* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ gamma, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ dres,
                                                           float* __restrict__ dx, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, LnGeom lg) {
  __shared__ float sred2[2][8];
  constexpr int KV = 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int vps = lg.C >> 2;
  const int vrow = vps * lg.nseg;
  const float inv_n = 1.0f / (float)(vrow * 4);
  float4 ag[KV], ab[KV], gm[KV];
#pragma unroll
  for (int k = 0; k < KV; ++k) {
    ag[k] = make_float4(0.f, 0.f, 0.f, 0.f); ab[k] = ag[k];
    const int v = tid + 256 * k;
    gm[k] = v < vrow ? __ldg(reinterpret_cast<const float4*>(gamma) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float* rsrc = dres != nullptr ? dres : x;
  const float rflag = dres != nullptr ? 1.0f : 0.0f;
  for (int row = blockIdx.x; row < lg.rows; row += gridDim.x) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[KV], gd[KV], rr[KV];
    long long idx[KV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const int v = tid + 256 * k;
      const bool on = v < vrow;
      idx[k] = -1;
      xh[k] = gd[k] = rr[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (on) {
        const int q = fdiv(v, lg.dvps), off = v - q * vps;
        const int srow = lg.mode == 2 ? merge_src(lg, row, q) : row;
        const float4 d = Vec4IO<YT>::cvt(Vec4IO<YT>::ldraw(dy, (long long)row * vrow + v));
        if (srow >= 0) {
          idx[k] = (long long)srow * vps + (lg.mode == 2 ? off : v);
          const float4 xv = Vec4IO<float>::ld(x, idx[k]);
          const float4 r4 = Vec4IO<float>::ld(rsrc, idx[k]);
          rr[k] = make_float4(r4.x * rflag, r4.y * rflag, r4.z * rflag, r4.w * rflag);
          xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        } else {
          xh[k] = make_float4(-mu * rs, -mu * rs, -mu * rs, -mu * rs);          // zero-padded segment of the merged row: xhat = (0 - mu) * rs
        }
        gd[k] = make_float4(d.x * gm[k].x, d.y * gm[k].y, d.z * gm[k].z, d.w * gm[k].w);
        s1 += gd[k].x + gd[k].y + gd[k].z + gd[k].w;
        s2 += gd[k].x * xh[k].x + gd[k].y * xh[k].y + gd[k].z * xh[k].z + gd[k].w * xh[k].w;
        ag[k].x += d.x * xh[k].x; ag[k].y += d.y * xh[k].y; ag[k].z += d.z * xh[k].z; ag[k].w += d.w * xh[k].w;
        ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
      }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { sred2[0][warp] = s1; sred2[1][warp] = s2; }
    __syncthreads();
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { m1 += sred2[0][w]; m2 += sred2[1][w]; }
    m1 *= inv_n; m2 *= inv_n;
    __syncthreads();                                   // sred2 is rewritten by the next row
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      if (idx[k] >= 0) {
        float4 o;
        o.x = rs * (gd[k].x - m1 - xh[k].x * m2) + rr[k].x;
        o.y = rs * (gd[k].y - m1 - xh[k].y * m2) + rr[k].y;
        o.z = rs * (gd[k].z - m1 - xh[k].z * m2) + rr[k].z;
        o.w = rs * (gd[k].w - m1 - xh[k].w * m2) + rr[k].w;
        Vec4IO<float>::st(dx, idx[k], o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KV; ++k) {
    const int v = tid + 256 * k;
    if (v < vrow) {
      atomicAdd(dgamma + v * 4 + 0, ag[k].x); atomicAdd(dgamma + v * 4 + 1, ag[k].y);
      atomicAdd(dgamma + v * 4 + 2, ag[k].z); atomicAdd(dgamma + v * 4 + 3, ag[k].w);
      atomicAdd(dbeta + v * 4 + 0, ab[k].x); atomicAdd(dbeta + v * 4 + 1, ab[k].y);
      atomicAdd(dbeta + v * 4 + 2, ab[k].z); atomicAdd(dbeta + v * 4 + 3, ab[k].w);
    }
  }
}

static void ln_shape(int vrow, int* G, int* vpl) {
  int g = 8;
  while (g < 32 && vrow > g * 4) g *= 2;
  *G = g;
  *vpl = ceil_div(vrow, g);
}

#define LN_CASES(X)                                                                                     \
  X(1, 8) X(2, 8) X(3, 8) X(4, 8) X(3, 16) X(4, 16) X(3, 32) X(4, 32) X(6, 32) X(8, 32) X(12, 32) X(16, 32) \
  X(24, 32) X(32, 32)

template <typename YT>
static int ln_fwd_dispatch(const swin_ln_args* a, const LnGeom& lg, cudaStream_t st) {
  int vrow = lg.C / 4 * lg.nseg, G, vpl;
  ln_shape(vrow, &G, &vpl);
  const int rows_per_block = 8 * (32 / G);
  long long blocks = ceil_div64(lg.rows, rows_per_block);
  int grid = (int)(blocks < (long long)kNumSMs * 8 ? blocks : (long long)kNumSMs * 8);
#define LN_FWD_CASE(V, GG)                                                                                    \
  if (G == GG && vpl <= V) {                                                                                  \
    ln_fwd_kernel<V, GG, YT><<<grid, 256, 0, st>>>(a->x, a->gamma, a->beta, (YT*)a->y, a->mean, a->rstd, lg); \
    SWIN_LAUNCH_CHECK();                                                                                      \
    return 0;                                                                                                 \
  }
  LN_CASES(LN_FWD_CASE)
#undef LN_FWD_CASE
  set_error("ln: row width %d too large", vrow * 4);
  return -EINVAL;
}

template <typename YT>
static int ln_bwd_dispatch(const swin_ln_args* a, const LnGeom& lg, cudaStream_t st) {
  int vrow = lg.C / 4 * lg.nseg, G, vpl;
  ln_shape(vrow, &G, &vpl);
  const int rows_per_block = 8 * (32 / G);
  // each lane group walks ~16 rows so the closing atomics amortise, but never fewer than 2 blocks per SM if rows allow
  long long blocks = ceil_div64(lg.rows, (long long)rows_per_block * 16);
  const long long min_blocks = ceil_div64(lg.rows, rows_per_block) < 2LL * kNumSMs ? ceil_div64(lg.rows, rows_per_block) : 2LL * kNumSMs;
  if (blocks < min_blocks) blocks = min_blocks;
  int grid = (int)(blocks < (long long)kNumSMs * 8 ? blocks : (long long)kNumSMs * 8);
  if (grid < 1) grid = 1;
  size_t smem = (size_t)2 * vrow * 4 * sizeof(float);
  if (lg.mode == 2 && vrow >= 256 && vrow <= 512 && a->dy2 == nullptr) {      // wide PatchMerging rows: one block per row
    const int gridb = lg.rows < 3 * kNumSMs ? lg.rows : 3 * kNumSMs;
    ln_bwd_block_kernel<YT><<<gridb, 256, 0, st>>>((const YT*)a->dy, a->x, a->gamma, a->mean, a->rstd, a->dres, a->dx, a->dgamma, a->dbeta, lg);
    SWIN_LAUNCH_CHECK();
    return 0;
  }
  static const bool use_pf = getenv("SWIN_LN_BWD_NO_PREFETCH") == nullptr;
#define LN_BWD_PF_CASE(V, GG)                                                                                      \
  if (use_pf && G == GG && vpl <= V) {                                                                             \
    const size_t smem_pf = ((smem + 15) & ~(size_t)15) + (size_t)8 * 2 * LnBwdStage<V, YT>::kBytes;                \
    { const int ar = ensure_dyn_smem((const void*)ln_bwd_pf_kernel<V, GG, YT>, 210 * 1024); if (ar) return ar; }      \
    ln_bwd_pf_kernel<V, GG, YT><<<grid, 256, smem_pf, st>>>((const YT*)a->dy, a->x, a->gamma, a->mean, a->rstd, a->dres, \
                                                           a->dx, a->dgamma, a->dbeta, lg);                        \
    SWIN_LAUNCH_CHECK();                                                                                           \
    return 0;                                                                                                      \
  }
  LN_BWD_PF_CASE(3, 8) LN_BWD_PF_CASE(4, 8) LN_BWD_PF_CASE(3, 16) LN_BWD_PF_CASE(4, 16) LN_BWD_PF_CASE(3, 32) LN_BWD_PF_CASE(4, 32)
  LN_BWD_PF_CASE(6, 32) LN_BWD_PF_CASE(8, 32)   // 768- / 1024-wide rows (stage-3 LayerNorms): 123-197 KB of staging, one block per SM either way
#undef LN_BWD_PF_CASE
#define LN_BWD_CASE(V, GG)                                                                                         \
  if (G == GG && vpl <= V) {                                                                                       \
    ln_bwd_kernel<V, GG, YT><<<grid, 256, smem, st>>>((const YT*)a->dy, a->x, a->gamma, a->mean, a->rstd, a->dres, \
                                                      a->dx, a->dgamma, a->dbeta, lg);                             \
    SWIN_LAUNCH_CHECK();                                                                                           \
    return 0;                                                                                                      \
  }
  LN_CASES(LN_BWD_CASE)
#undef LN_BWD_CASE
  set_error("ln: row width %d too large", vrow * 4);
  return -EINVAL;
}

// ------------------------------------------------------------------------------------------
// scale + cast (+ optional gather into window slots): dY of the residual epilogues.
// ------------------------------------------------------------------------------------------
template <typename YT, int kMaxV>
__global__ void __launch_bounds__(256) scale_cast_kernel(const float* __restrict__ x, YT* __restrict__ y,
                                                         const float* __restrict__ row_scale, int mode, WinGeom g,
                                                         int rows, float* __restrict__ colsum) {
  extern __shared__ float scol[];           // [C] block partial column sums (only when colsum != nullptr)
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int vrow = g.C >> 2;
  const int per_img_tok = g.H * g.W;
  float4 acc[kMaxV];
#pragma unroll
  for (int k = 0; k < kMaxV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (colsum != nullptr) {
    for (int i = threadIdx.x; i < g.C; i += blockDim.x) scol[i] = 0.f;
    __syncthreads();
  }
  // 4 rows per warp iteration: their loads are independent, so 4x the bytes are in flight per warp
  constexpr int RU = 4;
  for (int row0 = (blockIdx.x * wpb + (threadIdx.x >> 5)) * RU; row0 < rows; row0 += gridDim.x * wpb * RU) {
    int srow[RU];
    float sc[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int row = row0 + u;
      srow[u] = -2; sc[u] = 1.0f;
      if (row < rows) {
        int b, in;
        if (mode == 1) {
          b = split_slot_row(g, row, &in);
          int t = slot_to_token(g, in);
          srow[u] = t < 0 ? -1 : b * per_img_tok + t;
        } else {
          b = split_tok_row(g, row, &in);
          srow[u] = row;
        }
        if (row_scale) sc[u] = row_scale[b];
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxV; ++k) {
      const int v = lane + 32 * k;
      if (v < vrow) {
        float4 t[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) t[u] = srow[u] >= 0 ? Vec4IO<float>::ld(x, (long long)(srow[u] * vrow + v)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          if (srow[u] == -2) continue;
          t[u] = make_float4(t[u].x * sc[u], t[u].y * sc[u], t[u].z * sc[u], t[u].w * sc[u]);
          Vec4IO<YT>::st(y, (long long)((row0 + u) * vrow + v), t[u]);
          acc[k].x += t[u].x; acc[k].y += t[u].y; acc[k].z += t[u].z; acc[k].w += t[u].w;
        }
      }
    }
  }
  if (colsum != nullptr) {
#pragma unroll
    for (int k = 0; k < kMaxV; ++k) {
      const int v = lane + 32 * k;
      if (v < vrow) {
        atomicAdd(&scol[4 * v + 0], acc[k].x); atomicAdd(&scol[4 * v + 1], acc[k].y);
        atomicAdd(&scol[4 * v + 2], acc[k].z); atomicAdd(&scol[4 * v + 3], acc[k].w);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < g.C; i += blockDim.x) atomicAdd(colsum + i, scol[i]);
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4,
                                                        long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = i; v < n4; v += stride) Vec4IO<__nv_bfloat16>::st(y, v, Vec4IO<float>::ld(x, v));
  for (long long e = n4 * 4 + i; e < n; e += stride) y[e] = __float2bfloat16(x[e]);
}

// ------------------------------------------------------------------------------------------
// gradient gather: up to SWIN_GATHER_MAX fp32 tensors copied into their slots of one flat bucket by ONE launch
// (the data-parallel all-reduce bucket, ddp.py).  Block b serves the chunk (entry, 4096-float piece) found by scanning
// the per-entry cumulative chunk counts passed by value.
// ------------------------------------------------------------------------------------------
struct GatherTable {
  const float* src[SWIN_GATHER_MAX];
  long long dst_off[SWIN_GATHER_MAX];     // float offset inside the bucket (multiple of 4)
  int chunk_end[SWIN_GATHER_MAX];         // cumulative number of 4096-float chunks up to and including entry e
  int numel[SWIN_GATHER_MAX];
  int n;
};
constexpr int kGatherChunk = 4096;

__global__ void __launch_bounds__(256) grad_gather_kernel(const __grid_constant__ GatherTable t, float* __restrict__ bucket) {
  const int total = t.chunk_end[t.n - 1];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    int e = 0;
    while (c >= t.chunk_end[e]) ++e;
    const int first = e == 0 ? 0 : t.chunk_end[e - 1];
    const long long base = (long long)(c - first) * kGatherChunk;
    const int n = min(kGatherChunk, (int)(t.numel[e] - base));
    const float* __restrict__ src = t.src[e] + base;
    float* __restrict__ dst = bucket + t.dst_off[e] + base;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const int n4 = n >> 2;
      for (int v = threadIdx.x; v < n4; v += blockDim.x) reinterpret_cast<float4*>(dst)[v] = __ldg(reinterpret_cast<const float4*>(src) + v);
      for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused multi-tensor AdamW (decoupled weight decay, bias-corrected moments: torch.optim.AdamW arithmetic), one launch for
// up to SWIN_GATHER_MAX parameter tensors; optionally refreshes the bf16 shadow copy the GEMMs read, so the operand cast
// of the next step costs nothing.  Same chunk table scheme as the gradient gather.
// ------------------------------------------------------------------------------------------
struct AdamWTable {
  float* param[SWIN_GATHER_MAX];
  const float* grad[SWIN_GATHER_MAX];
  float* m[SWIN_GATHER_MAX];
  float* v[SWIN_GATHER_MAX];
  __nv_bfloat16* w16[SWIN_GATHER_MAX];
  float decay[SWIN_GATHER_MAX];           // 1 - lr * weight_decay of the tensor
  int chunk_end[SWIN_GATHER_MAX];
  int numel[SWIN_GATHER_MAX];
  int n;
  float beta1, beta2, omb1, omb2, eps, step_size, inv_sqrt_bc2, grad_scale;   // omb = 1 - beta, rounded from double like torch's scalars
};

__global__ void __launch_bounds__(256) adamw_kernel(const __grid_constant__ AdamWTable t) {
  const int total = t.chunk_end[t.n - 1];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    int e = 0;
    while (c >= t.chunk_end[e]) ++e;
    const int first = e == 0 ? 0 : t.chunk_end[e - 1];
    const long long base = (long long)(c - first) * kGatherChunk;
    const int n = min(kGatherChunk, (int)(t.numel[e] - base));
    float* __restrict__ pp = t.param[e] + base;
    const float* __restrict__ gg = t.grad[e] + base;
    float* __restrict__ mm = t.m[e] + base;
    float* __restrict__ vv = t.v[e] + base;
    __nv_bfloat16* __restrict__ ww = t.w16[e] ? t.w16[e] + base : nullptr;
    const float decay = t.decay[e];
    auto update = [&](float g, float& pv, float& m, float& v) {
      g *= t.grad_scale;
      pv *= decay;
      m = m + (g - m) * t.omb1;                  // exp_avg.lerp_(grad, 1 - beta1)
      v = v * t.beta2 + t.omb2 * g * g;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(v) * t.inv_sqrt_bc2 + t.eps;
      pv = pv - t.step_size * (m / denom);
    };
    // 128-bit accesses (the chunk base is a multiple of 4 elements; torch's allocations are 512-byte aligned, so only views at
    // odd offsets take the scalar path); four independent 16-byte loads per thread are in flight before the first use
    const bool vec = ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(gg) | reinterpret_cast<uintptr_t>(mm) |
                       reinterpret_cast<uintptr_t>(vv)) & 15) == 0 && (reinterpret_cast<uintptr_t>(ww) & 7) == 0;
    const int n4 = vec ? n >> 2 : 0;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gg) + i);
      float4 p4 = reinterpret_cast<float4*>(pp)[i], m4 = reinterpret_cast<float4*>(mm)[i], v4 = reinterpret_cast<float4*>(vv)[i];
      update(g4.x, p4.x, m4.x, v4.x); update(g4.y, p4.y, m4.y, v4.y); update(g4.z, p4.z, m4.z, v4.z); update(g4.w, p4.w, m4.w, v4.w);
      reinterpret_cast<float4*>(pp)[i] = p4; reinterpret_cast<float4*>(mm)[i] = m4; reinterpret_cast<float4*>(vv)[i] = v4;
      if (ww) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(ww)[i] = pk;
      }
    }
    for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) {
      float pv = pp[i], m = mm[i], v = vv[i];
      update(gg[i], pv, m, v);
      pp[i] = pv; mm[i] = m; vv[i] = v;
      if (ww) ww[i] = __float2bfloat16(pv);
    }
  }
}

// ------------------------------------------------------------------------------------------
// column sums (bias gradients): block = 256 threads = 8 row-lanes x 32 column-vectors(4 wide)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, int M, int N, long long ld, float* __restrict__ out,
                                                     int rows_per_block) {
  __shared__ float4 part[8][32];
  const int cv = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cv) * 4;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) {
    for (int r = r0 + rl; r < r1; r += 8) {
      float4 v = Vec4IO<T>::ld(X + (long long)r * ld + col, 0);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  part[rl][cv] = acc;
  __syncthreads();
  if (rl == 0 && col < N) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { float4 v = part[k][cv]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    atomicAdd(out + col + 0, acc.x); atomicAdd(out + col + 1, acc.y);
    atomicAdd(out + col + 2, acc.z); atomicAdd(out + col + 3, acc.w);
  }
}

}  // namespace swin

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace swin;

extern "C" int swin_window_partition(const void* x, void* win, int B, int Hp, int Wp, int C, int ws, int elem_bytes, void* stream) {
  SWIN_REQUIRE(ws > 0 && Hp % ws == 0 && Wp % ws == 0, "window_partition: Hp,Wp must be multiples of ws");
  WinGeom g = make_geom(B, Hp, Wp, C, ws, 0);
  return launch_row_map<MAP_PARTITION>(x, win, g, elem_bytes, (long long)B * g.nW * g.N, (cudaStream_t)stream);
}
extern "C" int swin_window_reverse(const void* win, void* x, int B, int Hp, int Wp, int C, int ws, int elem_bytes, void* stream) {
  SWIN_REQUIRE(ws > 0 && Hp % ws == 0 && Wp % ws == 0, "window_reverse: Hp,Wp must be multiples of ws");
  WinGeom g = make_geom(B, Hp, Wp, C, ws, 0);
  return launch_row_map<MAP_REVERSE>(win, x, g, elem_bytes, (long long)B * Hp * Wp, (cudaStream_t)stream);
}
extern "C" int swin_window_gather(const void* x, void* xw, int B, int H, int W, int C, int ws, int shift, int elem_bytes, void* stream) {
  SWIN_REQUIRE(ws > 0 && shift >= 0 && shift < ws, "window_gather: bad ws/shift");
  WinGeom g = make_geom(B, H, W, C, ws, shift);
  return launch_row_map<MAP_GATHER>(x, xw, g, elem_bytes, (long long)B * g.nW * g.N, (cudaStream_t)stream);
}
extern "C" int swin_window_scatter(const void* xw, void* x, int B, int H, int W, int C, int ws, int shift, int elem_bytes, void* stream) {
  SWIN_REQUIRE(ws > 0 && shift >= 0 && shift < ws, "window_scatter: bad ws/shift");
  WinGeom g = make_geom(B, H, W, C, ws, shift);
  return launch_row_map<MAP_SCATTER>(xw, x, g, elem_bytes, (long long)B * H * W, (cudaStream_t)stream);
}
extern "C" int swin_shift_mask(float* mask, int H, int W, int ws, int shift, void* stream) {
  SWIN_REQUIRE(H > 0 && W > 0 && ws > 0 && shift >= 0 && shift < ws, "shift_mask: bad geometry");
  WinGeom g = make_geom(1, H, W, 1, ws, shift);
  long long total = (long long)g.nW * g.N * g.N;
  int grid = (int)((total + 255) / 256 < kNumSMs * 8 ? (total + 255) / 256 : kNumSMs * 8);
  shift_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mask, g);
  SWIN_LAUNCH_CHECK();
  return 0;
}
extern "C" int swin_mask_nonzero(const float* mask, int32_t* flags, int nW, int N, void* stream) {
  SWIN_REQUIRE(mask && flags && nW > 0 && N > 0, "mask_nonzero: bad arguments");
  mask_nonzero_kernel<<<nW, 256, 0, (cudaStream_t)stream>>>(mask, flags, nW, N * N);
  SWIN_LAUNCH_CHECK();
  return 0;
}
extern "C" int swin_rel_bias_expand(const float* table, float* bias, int nH, int ws, void* stream) {
  SWIN_REQUIRE(nH > 0 && ws > 0, "rel_bias_expand: bad shape");
  int total = nH * ws * ws * ws * ws;
  rel_bias_expand_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(table, bias, nH, ws);
  SWIN_LAUNCH_CHECK();
  return 0;
}
extern "C" int swin_rel_bias_reduce(const float* dbias, float* dtable, int nH, int ws, void* stream) {
  SWIN_REQUIRE(nH > 0 && ws > 0, "rel_bias_reduce: bad shape");
  int total = (2 * ws - 1) * (2 * ws - 1) * nH;
  rel_bias_reduce_kernel<<<ceil_div(total, 8), 256, 0, (cudaStream_t)stream>>>(dbias, dtable, nH, ws);
  SWIN_LAUNCH_CHECK();
  return 0;
}

extern "C" int swin_ln_fwd(const swin_ln_args* a, void* stream) {
  LnGeom lg;
  int rc = ln_geom(a, false, &lg);
  if (rc) return rc;
  SWIN_REQUIRE(a->x && a->gamma && a->beta && a->y && a->mean && a->rstd, "ln_fwd: null pointer");
  SWIN_REQUIRE(aligned16(a->x) && aligned16(a->y) && aligned16(a->gamma) && aligned16(a->beta), "ln_fwd: alignment");
  if (lg.rows == 0) return 0;
  if (a->y_dtype == SWIN_F32) return ln_fwd_dispatch<float>(a, lg, (cudaStream_t)stream);
  return ln_fwd_dispatch<__nv_bfloat16>(a, lg, (cudaStream_t)stream);
}
extern "C" int swin_ln_bwd(const swin_ln_args* a, void* stream) {
  LnGeom lg;
  int rc = ln_geom(a, true, &lg);
  if (rc) return rc;
  SWIN_REQUIRE(a->x && a->gamma && a->dy && a->mean && a->rstd && a->dx && a->dgamma && a->dbeta, "ln_bwd: null pointer");
  SWIN_REQUIRE(aligned16(a->x) && aligned16(a->dy) && aligned16(a->dx) && aligned16(a->gamma), "ln_bwd: alignment");
  SWIN_REQUIRE(a->dres == nullptr || aligned16(a->dres), "ln_bwd: alignment");
  if (lg.rows == 0) return 0;
  if (a->y_dtype == SWIN_F32) return ln_bwd_dispatch<float>(a, lg, (cudaStream_t)stream);
  return ln_bwd_dispatch<__nv_bfloat16>(a, lg, (cudaStream_t)stream);
}

extern "C" int swin_scale_cast(const float* x, void* y, const float* row_scale, int mode, int B, int H, int W, int C, int ws,
                               int shift, int y_dtype, float* colsum, void* stream) {
  SWIN_REQUIRE(mode == 0 || mode == 1, "scale_cast: bad mode");
  SWIN_REQUIRE(C % 4 == 0 && C <= 1024 && B > 0 && H > 0 && W > 0, "scale_cast: bad shape (C %% 4 == 0, C <= 1024)");
  SWIN_REQUIRE(aligned16(x) && aligned16(y), "scale_cast: alignment");
  if (mode == 0) { ws = 1; shift = 0; }
  SWIN_REQUIRE(ws > 0 && shift >= 0 && shift < ws, "scale_cast: bad window geometry");
  WinGeom g = make_geom(B, H, W, C, ws, shift);
  long long rows = mode == 1 ? (long long)B * g.nW * g.N : (long long)B * H * W;
  SWIN_REQUIRE(rows < (1ll << 30) && rows * (C / 4) < (1ll << 31), "scale_cast: tensor too large for 32-bit row / vector indices");
  long long blocks = ceil_div64(rows, 8 * 4 * 4);
  int grid = (int)(blocks < (long long)kNumSMs * 8 ? blocks : (long long)kNumSMs * 8);
  if (grid < 1) grid = 1;
  size_t smem = colsum ? (size_t)C * sizeof(float) : 0;
  const int kv = ceil_div(C / 4, 32);           // float4 vectors per lane (C <= 1024 -> <= 8)
#define SC_LAUNCH(T, KV) scale_cast_kernel<T, KV><<<grid, 256, smem, (cudaStream_t)stream>>>(x, (T*)y, row_scale, mode, g, (int)rows, colsum)
#define SC_DISPATCH(T)                                                                          \
  do {                                                                                          \
    if (kv <= 1) SC_LAUNCH(T, 1); else if (kv <= 2) SC_LAUNCH(T, 2); else if (kv <= 3) SC_LAUNCH(T, 3); \
    else if (kv <= 4) SC_LAUNCH(T, 4); else if (kv <= 6) SC_LAUNCH(T, 6); else SC_LAUNCH(T, 8);   \
  } while (0)
  if (y_dtype == SWIN_F32) SC_DISPATCH(float);
  else if (y_dtype == SWIN_BF16) SC_DISPATCH(__nv_bfloat16);
  else { set_error("scale_cast: bad dtype"); return -EINVAL; }
#undef SC_DISPATCH
#undef SC_LAUNCH
  SWIN_LAUNCH_CHECK();
  return 0;
}
extern "C" int swin_cast_bf16(const float* x, void* y, int64_t n, void* stream) {
  SWIN_REQUIRE(n >= 0 && aligned16(x) && aligned16(y), "cast_bf16: alignment");
  if (n == 0) return 0;
  long long n4 = n / 4;
  long long blocks = ceil_div64(n4 > 0 ? n4 : 1, 256);
  int grid = (int)(blocks < (long long)kNumSMs * 8 ? blocks : (long long)kNumSMs * 8);
  cast_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y, n4, n);
  SWIN_LAUNCH_CHECK();
  return 0;
}
extern "C" int swin_grad_gather(const void* const* src, const int64_t* dst_off, const int64_t* numel, int n, float* bucket, void* stream) {
  SWIN_REQUIRE(n >= 0 && n <= SWIN_GATHER_MAX, "grad_gather: at most %d tensors per call (got %d)", SWIN_GATHER_MAX, n);
  if (n == 0) return 0;
  SWIN_REQUIRE(src && dst_off && numel && bucket && aligned16(bucket), "grad_gather: null / misaligned pointer");
  GatherTable t;
  int chunks = 0, live = 0;
  for (int e = 0; e < n; ++e) {
    SWIN_REQUIRE(numel[e] >= 0 && numel[e] < (1ll << 31) && dst_off[e] >= 0 && dst_off[e] % 4 == 0, "grad_gather: bad entry %d", e);
    if (numel[e] == 0) continue;
    SWIN_REQUIRE(src[e] != nullptr && (reinterpret_cast<uintptr_t>(src[e]) & 3) == 0, "grad_gather: entry %d null / misaligned", e);
    chunks += (int)ceil_div64(numel[e], kGatherChunk);
    t.src[live] = (const float*)src[e]; t.dst_off[live] = dst_off[e]; t.numel[live] = (int)numel[e]; t.chunk_end[live] = chunks;
    ++live;
  }
  if (live == 0) return 0;
  t.n = live;
  const int grid = chunks < kNumSMs * 8 ? chunks : kNumSMs * 8;
  grad_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, bucket);
  SWIN_LAUNCH_CHECK();
  return 0;
}
extern "C" int swin_adamw_step(void* const* param, const void* const* grad, void* const* exp_avg, void* const* exp_avg_sq,
                               void* const* w16, const float* weight_decay, const int64_t* numel, int n, double lr, double beta1,
                               double beta2, double eps, int step, double grad_scale, void* stream) {
  SWIN_REQUIRE(n >= 0 && n <= SWIN_GATHER_MAX, "adamw: at most %d tensors per call (got %d)", SWIN_GATHER_MAX, n);
  SWIN_REQUIRE(step >= 1 && lr >= 0.0 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "adamw: bad hyper-parameters");
  if (n == 0) return 0;
  SWIN_REQUIRE(param && grad && exp_avg && exp_avg_sq && weight_decay && numel, "adamw: null table");
  AdamWTable t;
  int chunks = 0, live = 0;
  for (int e = 0; e < n; ++e) {
    SWIN_REQUIRE(numel[e] >= 0 && numel[e] < (1ll << 31), "adamw: bad numel in entry %d", e);
    if (numel[e] == 0) continue;
    SWIN_REQUIRE(param[e] && grad[e] && exp_avg[e] && exp_avg_sq[e], "adamw: null tensor in entry %d", e);
    chunks += (int)ceil_div64(numel[e], kGatherChunk);
    t.param[live] = (float*)param[e]; t.grad[live] = (const float*)grad[e]; t.m[live] = (float*)exp_avg[e]; t.v[live] = (float*)exp_avg_sq[e];
    t.w16[live] = w16 ? (__nv_bfloat16*)w16[e] : nullptr;
    t.decay[live] = (float)(1.0 - lr * (double)weight_decay[e]);
    t.numel[live] = (int)numel[e]; t.chunk_end[live] = chunks;
    ++live;
  }
  if (live == 0) return 0;
  t.n = live;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  t.beta1 = (float)beta1; t.beta2 = (float)beta2; t.eps = (float)eps; t.grad_scale = (float)grad_scale;
  t.omb1 = (float)(1.0 - beta1); t.omb2 = (float)(1.0 - beta2);
  t.step_size = (float)(lr / bc1);
  t.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const int grid = chunks < kNumSMs * 8 ? chunks : kNumSMs * 8;
  adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t);
  SWIN_LAUNCH_CHECK();
  return 0;
}
extern "C" int swin_colsum(const void* X, int M, int N, int64_t ld, int dtype, float* colsum, void* stream) {
  SWIN_REQUIRE(M >= 0 && N > 0 && N % 4 == 0 && ld >= N && ld % 4 == 0, "colsum: bad shape");
  SWIN_REQUIRE(aligned16(X) && aligned16(colsum), "colsum: alignment");
  if (M == 0) return 0;
  int gx = ceil_div(N, 128);
  int target_y = ceil_div(kNumSMs * 4, gx);
  int rpb = ceil_div(M, target_y);
  if (rpb < 64) rpb = 64;
  dim3 grid(gx, ceil_div(M, rpb));
  if (dtype == SWIN_F32) colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)X, M, N, ld, colsum, rpb);
  else if (dtype == SWIN_BF16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)X, M, N, ld, colsum, rpb);
  else { set_error("colsum: bad dtype"); return -EINVAL; }
  SWIN_LAUNCH_CHECK();
  return 0;
}
