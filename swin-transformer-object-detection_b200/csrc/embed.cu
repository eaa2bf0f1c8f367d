// The two ends of the block stack:
//   * PatchEmbed (REF:429-445): zero-pad + 4x4/4 conv as a patch-unfold gather feeding the GEMM kernels;
//   * per-stage output norm + NHWC->NCHW (REF:618-623): LayerNorm whose store is transposed through smem,
//     and the matching backward that reads the NCHW gradient through the same transpose.
#include "common.cuh"

namespace swin {

// ------------------------------------------------------------------------------------------ patch unfold
// img (B, Cin, Hi, Wi) fp32  <->  cols (B*Hh*Ww, Cin*p*p), column order [c][i][j] == weight.view(C, -1)
template <typename T, bool SCATTER>
__global__ void __launch_bounds__(256) patch_unfold_kernel(float* __restrict__ img, T* __restrict__ cols, int B, int Cin, int Hi,
                                                           int Wi, int p, int Hh, int Ww) {
  const long long total = (long long)B * Cin * Hh * p * Ww;
  const int K = Cin * p * p;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(e % Ww);
    long long r = e / Ww;
    const int y = (int)(r % (Hh * p)); r /= (Hh * p);
    const int c = (int)(r % Cin);
    const int b = (int)(r / Cin);
    const int ph = y / p, i = y - ph * p;
    T* dst = cols + ((long long)(b * Hh + ph) * Ww + pw) * K + c * p * p + i * p;
    float* src = img + (((long long)b * Cin + c) * Hi + y) * Wi + (long long)pw * p;
    if (!SCATTER && p == 4) {                  // the 4 pixels of one patch row -> one 8/16-byte store
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = ((y < Hi) && (pw * 4 + j < Wi)) ? __ldg(src + j) : 0.f;
      if (sizeof(T) == 2) *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
      else *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      continue;
    }
    for (int j = 0; j < p; ++j) {
      const bool in = (y < Hi) && (pw * p + j < Wi);
      if (SCATTER) { if (in) src[j] = (float)dst[j]; }
      else dst[j] = (T)(in ? src[j] : 0.f);
    }
  }
}

// Forward gather for 4x4 patches and bf16 columns (the benchmark's PatchEmbed): one thread = one (token, channel): four 16-byte
// image-row reads (consecutive threads = consecutive patches of one image row: coalesced) and ONE whole 32-byte sector written
// ([c][i][j] = 16 bf16).  The generic kernel above writes 8-byte pieces 96 bytes apart (2.1 TB/s); this one is a plain stream.
__global__ void __launch_bounds__(256) patch_unfold4_bf16_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ cols, int Cin,
                                                                 int Hi, int Wi, int Hh, int Ww) {
  const int bp = blockIdx.y;                         // b * Hh + ph
  const int b = bp / Hh, ph = bp - b * Hh;
  const int K = Cin * 16;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Cin * Ww; t += gridDim.x * blockDim.x) {
    const int c = t / Ww, pw = t - c * Ww;
    const float* src = img + (((long long)b * Cin + c) * Hi + ph * 4) * Wi + pw * 4;
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v[4];
      const bool rowin = ph * 4 + i < Hi;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (rowin && pw * 4 + j < Wi) ? __ldg(src + (long long)i * Wi + j) : 0.f;
      o[2 * i] = pack_bf16(v[0], v[1]); o[2 * i + 1] = pack_bf16(v[2], v[3]);
    }
    uint4* dst = reinterpret_cast<uint4*>(cols + ((long long)bp * Ww + pw) * K + c * 16);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

// ------------------------------------------------------------------------------------------ LN + NCHW store
// A block owns a tile of 32 consecutive tokens.  Each token is normalised by a group of G lanes holding VPL float4
// (like the LN kernels in elementwise.cu), so 8 warps x 32/G tokens are in flight per round; the tile is then
// written (read) channel-major so that global accesses are 128-byte rows of 32 tokens.
constexpr int kTokTile = 32;

template <int G>
__device__ __forceinline__ float group_sum_e(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int VPL, int G, int TOK>
__global__ void __launch_bounds__(256) ln_nchw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ out,
                                                          float* __restrict__ mean, float* __restrict__ rstd, int L, int C, float eps) {
  extern __shared__ float tile[];            // [C][TOK + 1]
  constexpr int pitch = TOK + 1;
  constexpr int R = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane % G, gi = lane / G;
  const int b = blockIdx.y, t0 = blockIdx.x * TOK;
  const int vrow = C >> 2;
  const float inv_n = 1.0f / (float)C;
#pragma unroll 1
  for (int tt = warp * R + gi; tt < TOK; tt += 8 * R) {
    const int t = t0 + tt;
    const bool act = t < L;
    const float4* row = reinterpret_cast<const float4*>(x + ((long long)b * L + (act ? t : 0)) * C);
    float4 r[VPL];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = gl + G * k;
      r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act && v < vrow) { r[k] = __ldg(row + v); s += r[k].x + r[k].y + r[k].z + r[k].w; }
    }
    const float mu = group_sum_e<G>(s) * inv_n;
    float q2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
      if (gl + G * k < vrow) {
        float a = r[k].x - mu, b2 = r[k].y - mu, c2 = r[k].z - mu, d = r[k].w - mu;
        q2 += a * a + b2 * b2 + c2 * c2 + d * d;
      }
    const float rs = rsqrtf(group_sum_e<G>(q2) * inv_n + eps);
    if (act && gl == 0) { mean[(long long)b * L + t] = mu; rstd[(long long)b * L + t] = rs; }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = gl + G * k;
      if (act && v < vrow) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + v), bt = __ldg(reinterpret_cast<const float4*>(beta) + v);
        const int c = 4 * v;
        tile[(c + 0) * pitch + tt] = (r[k].x - mu) * rs * gm.x + bt.x;
        tile[(c + 1) * pitch + tt] = (r[k].y - mu) * rs * gm.y + bt.y;
        tile[(c + 2) * pitch + tt] = (r[k].z - mu) * rs * gm.z + bt.z;
        tile[(c + 3) * pitch + tt] = (r[k].w - mu) * rs * gm.w + bt.w;
      }
    }
  }
  __syncthreads();
  constexpr int cper = 32 / TOK;             // a warp stores cper channels x TOK tokens per instruction
  const int t = t0 + (lane & (TOK - 1)), cl = lane / TOK;
  if (t < L)
    for (int c = warp * cper + cl; c < C; c += 8 * cper) out[((long long)b * C + c) * L + t] = tile[c * pitch + (lane & (TOK - 1))];
}

template <int VPL, int G, int TOK, bool SX>
__global__ void __launch_bounds__(256, (VPL <= 3 ? 3 : (VPL <= 6 ? 2 : 1))) ln_nchw_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                          const float* __restrict__ gamma, const float* __restrict__ mean,
                                                          const float* __restrict__ rstd, float* __restrict__ dx,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, int L, int C,
                                                          int tiles_per_block, int nbuf) {
  constexpr int tok = TOK;
  constexpr bool stage_x = SX;
  // dout arrives channel-major (NCHW): a [C][tok-token] tile (tok = 32, or 16 for wide rows) is transposed through smem.
  // With two buffers the NEXT tile's dout (and, with stage_x, its x rows and statistics) is fetched with cp.async while the
  // current one is processed.  Without stage_x the rows of x are read straight from global memory (float4 per lane, all
  // VPL loads of a token issued together).
  extern __shared__ __align__(16) float tile_all[];        // nbuf x ([C][tok+1] dout tile (+ [tok][C] x rows + 2 x tok stats)) + [2][C] partials
  constexpr int pitch = tok + 1;
  const size_t buf_floats = (size_t)C * pitch + (stage_x ? (size_t)tok * C + 2 * tok : 0) + 4;      // +4 keeps the x rows 16-byte aligned
  float* sred = tile_all + (size_t)nbuf * buf_floats;
  constexpr int R = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane % G, gi = lane / G;
  const int b = blockIdx.y;
  const int vrow = C >> 2;
  const float inv_n = 1.0f / (float)C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;
  float4 ag[VPL], ab[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) { ag[k] = make_float4(0.f, 0.f, 0.f, 0.f); ab[k] = ag[k]; }
  const size_t xoff = ((size_t)C * pitch + 3) & ~(size_t)3;
  constexpr int cper = 32 / tok;
  const int tl = lane & (tok - 1), cl = lane / tok;      // a warp covers cper channels x tok tokens per load
  auto fetch_tile = [&](int tb, int buf) {
    const int t0 = (blockIdx.x * tiles_per_block + tb) * tok;
    if (tb >= tiles_per_block || t0 >= L) return;
    float* base = tile_all + (size_t)buf * buf_floats;
    {
      const int t = t0 + tl;
      const uint32_t nbytes = t < L ? 4u : 0u;           // src-size 0: zero fill
      const float* src = dout + (long long)b * C * L + (t < L ? t : 0);
      for (int c = warp * cper + cl; c < C; c += 8 * cper)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(base + c * pitch + tl)),
                     "l"(src + (long long)c * L), "r"(nbytes) : "memory");
    }
    if (stage_x) {
      // x rows of the tile: tok tokens x C floats, contiguous in global memory (token-major)
      const int nvec = tok * vrow;
      const int live = min(tok, L - t0) * vrow;            // vectors of real tokens (the tile's rows are contiguous)
      const long long row0 = (long long)b * L + t0;
      for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        const uint32_t nbytes = v < live ? 16u : 0u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(base + xoff + (size_t)v * 4)),
                     "l"(reinterpret_cast<const float4*>(x + row0 * C) + (v < live ? v : 0)), "r"(nbytes) : "memory");
      }
      // per-token statistics of the tile (mean | rstd), 2 x tok floats after the x rows
      if ((int)threadIdx.x < 2 * tok) {
        const int tt = threadIdx.x & (tok - 1);
        const float* sp = ((int)threadIdx.x < tok ? mean : rstd) + row0 + (t0 + tt < L ? tt : 0);
        const uint32_t nb = t0 + tt < L ? 4u : 0u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(base + xoff + (size_t)nvec * 4 + threadIdx.x)),
                     "l"(sp), "r"(nb) : "memory");
      }
    }
  };
  fetch_tile(0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int tb = 0; tb < tiles_per_block; ++tb) {
    const int t0 = (blockIdx.x * tiles_per_block + tb) * tok;
    if (t0 >= L) break;
    if (nbuf == 2) {
      fetch_tile(tb + 1, (tb + 1) & 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* tile = tile_all + (size_t)(nbuf == 2 ? (tb & 1) : 0) * buf_floats;
    const float4* xs = reinterpret_cast<const float4*>(tile + xoff);
#pragma unroll 1
    for (int tt = warp * R + gi; tt < tok; tt += 8 * R) {
      const int t = t0 + tt;
      const bool act = t < L;
      const long long rowi = (long long)b * L + (act ? t : 0);
      const float* stat = tile + xoff + (size_t)tok * vrow * 4;
      const float mu = stage_x ? stat[tt] : __ldg(mean + rowi), rs = stage_x ? stat[tok + tt] : __ldg(rstd + rowi);
      const float4* xrow = stage_x ? xs + (size_t)tt * vrow : reinterpret_cast<const float4*>(x + rowi * C);
      float4 xh[VPL], gd[VPL];
      if (!stage_x) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {           // the token's x vectors: every load in flight before the first use
          const int v = gl + G * k;
          xh[k] = (act && v < vrow) ? __ldg(xrow + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int v = gl + G * k;
        const float4 xv = stage_x ? ((act && v < vrow) ? xrow[v] : make_float4(0.f, 0.f, 0.f, 0.f)) : xh[k];
        xh[k] = make_float4(0.f, 0.f, 0.f, 0.f); gd[k] = xh[k];
        if (act && v < vrow) {
          const int c = 4 * v;
          const float4 d = make_float4(tile[(c + 0) * pitch + tt], tile[(c + 1) * pitch + tt], tile[(c + 2) * pitch + tt], tile[(c + 3) * pitch + tt]);
          const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + v);
          xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
          gd[k] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
          s1 += gd[k].x + gd[k].y + gd[k].z + gd[k].w;
          s2 += gd[k].x * xh[k].x + gd[k].y * xh[k].y + gd[k].z * xh[k].z + gd[k].w * xh[k].w;
          ag[k].x += d.x * xh[k].x; ag[k].y += d.y * xh[k].y; ag[k].z += d.z * xh[k].z; ag[k].w += d.w * xh[k].w;
          ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
        }
      }
      const float m1 = group_sum_e<G>(s1) * inv_n, m2 = group_sum_e<G>(s2) * inv_n;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int v = gl + G * k;
        if (act && v < vrow)
          reinterpret_cast<float4*>(dx + rowi * C)[v] = make_float4(rs * (gd[k].x - m1 - xh[k].x * m2), rs * (gd[k].y - m1 - xh[k].y * m2),
                                                                    rs * (gd[k].z - m1 - xh[k].z * m2), rs * (gd[k].w - m1 - xh[k].w * m2));
      }
    }
    __syncthreads();                          // every warp is done with this tile before its buffer is refilled
    if (nbuf == 1) {                          // single buffer (one or two tiles per block, overlap comes from the co-resident blocks)
      fetch_tile(tb + 1, 0);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int v = gl + G * k;
    if (v < vrow) {
      atomicAdd(&sred[4 * v + 0], ag[k].x); atomicAdd(&sred[4 * v + 1], ag[k].y);
      atomicAdd(&sred[4 * v + 2], ag[k].z); atomicAdd(&sred[4 * v + 3], ag[k].w);
      atomicAdd(&sred[C + 4 * v + 0], ab[k].x); atomicAdd(&sred[C + 4 * v + 1], ab[k].y);
      atomicAdd(&sred[C + 4 * v + 2], ab[k].z); atomicAdd(&sred[C + 4 * v + 3], ab[k].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) { atomicAdd(dgamma + i, sred[i]); atomicAdd(dbeta + i, sred[C + i]); }
}

static void nchw_shape(int vrow, int* G, int* vpl) {
  int g = 8;
  while (g < 32 && vrow > g * 4) g *= 2;
  *G = g;
  *vpl = (vrow + g - 1) / g;
}
#define NCHW_CASES(X) X(1, 8) X(2, 8) X(3, 8) X(4, 8) X(3, 16) X(4, 16) X(3, 32) X(4, 32) X(6, 32) X(8, 32)

}  // namespace swin

using namespace swin;

static int patch_geom_check(int B, int Cin, int Hi, int Wi, int p) {
  SWIN_REQUIRE(B > 0 && Cin > 0 && Hi > 0 && Wi > 0 && p > 0 && p <= 16, "patch: bad shape");
  return 0;
}

extern "C" int swin_patch_gather(const float* img, void* cols, int B, int Cin, int Hi, int Wi, int patch, int dtype, void* stream) {
  int rc = patch_geom_check(B, Cin, Hi, Wi, patch);
  if (rc) return rc;
  const int Hh = ceil_div(Hi, patch), Ww = ceil_div(Wi, patch);
  long long total = (long long)B * Cin * Hh * patch * Ww;
  int grid = (int)(ceil_div64(total, 256) < (long long)kNumSMs * 16 ? ceil_div64(total, 256) : (long long)kNumSMs * 16);
  if (dtype == SWIN_BF16 && patch == 4 && aligned16(cols)) {
    dim3 g4((unsigned)ceil_div(Cin * Ww, 256), (unsigned)(B * Hh));
    patch_unfold4_bf16_kernel<<<g4, 256, 0, (cudaStream_t)stream>>>(img, (__nv_bfloat16*)cols, Cin, Hi, Wi, Hh, Ww);
    SWIN_LAUNCH_CHECK();
    return 0;
  }
  if (dtype == SWIN_F32) patch_unfold_kernel<float, false><<<grid, 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(img), (float*)cols, B, Cin, Hi, Wi, patch, Hh, Ww);
  else if (dtype == SWIN_BF16) patch_unfold_kernel<__nv_bfloat16, false><<<grid, 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(img), (__nv_bfloat16*)cols, B, Cin, Hi, Wi, patch, Hh, Ww);
  else { set_error("patch_gather: bad dtype"); return -EINVAL; }
  SWIN_LAUNCH_CHECK();
  return 0;
}
extern "C" int swin_patch_scatter(const void* dcols, float* dimg, int B, int Cin, int Hi, int Wi, int patch, int dtype, void* stream) {
  int rc = patch_geom_check(B, Cin, Hi, Wi, patch);
  if (rc) return rc;
  const int Hh = ceil_div(Hi, patch), Ww = ceil_div(Wi, patch);
  long long total = (long long)B * Cin * Hh * patch * Ww;
  int grid = (int)(ceil_div64(total, 256) < (long long)kNumSMs * 16 ? ceil_div64(total, 256) : (long long)kNumSMs * 16);
  if (dtype == SWIN_F32) patch_unfold_kernel<float, true><<<grid, 256, 0, (cudaStream_t)stream>>>(dimg, (float*)const_cast<void*>(dcols), B, Cin, Hi, Wi, patch, Hh, Ww);
  else if (dtype == SWIN_BF16) patch_unfold_kernel<__nv_bfloat16, true><<<grid, 256, 0, (cudaStream_t)stream>>>(dimg, (__nv_bfloat16*)const_cast<void*>(dcols), B, Cin, Hi, Wi, patch, Hh, Ww);
  else { set_error("patch_scatter: bad dtype"); return -EINVAL; }
  SWIN_LAUNCH_CHECK();
  return 0;
}

extern "C" int swin_ln_nchw_fwd(const float* x, const float* gamma, const float* beta, float* out, float* mean, float* rstd, int B,
                                int L, int C, float eps, void* stream) {
  SWIN_REQUIRE(B > 0 && L > 0 && C > 0 && C % 4 == 0 && C <= 1024, "ln_nchw: bad shape (C %% 4 == 0, C <= 1024)");
  SWIN_REQUIRE(x && gamma && beta && out && mean && rstd, "ln_nchw: null pointer");
  SWIN_REQUIRE(aligned16(x) && aligned16(gamma) && aligned16(beta), "ln_nchw: alignment");
  // 32-token tiles (128-byte NCHW row segments) up to C = 256; wider rows come with short token axes, where 8-token tiles
  // (one 32-byte sector per channel row, a quarter of the shared memory, more blocks per SM) measured 8-35 % faster
  const int tok = C <= 256 ? kTokTile : kTokTile / 4;
  size_t smem = (size_t)C * (tok + 1) * sizeof(float);
  int G, vpl;
  nchw_shape(C / 4, &G, &vpl);
  dim3 grid(ceil_div(L, tok), B);
#define NCHW_FWD_LAUNCH(V, GG, TK)                                                                                        \
  if (tok == TK) {                                                                                                        \
    const int ar = ensure_dyn_smem((const void*)ln_nchw_fwd_kernel<V, GG, TK>, 160 * 1024);                               \
    if (ar) return ar;                                                                                                    \
    ln_nchw_fwd_kernel<V, GG, TK><<<grid, 256, smem, (cudaStream_t)stream>>>(x, gamma, beta, out, mean, rstd, L, C, eps); \
  }
#define NCHW_FWD(V, GG)                                                                                                   \
  if (G == GG && vpl <= V) {                                                                                              \
    NCHW_FWD_LAUNCH(V, GG, kTokTile) NCHW_FWD_LAUNCH(V, GG, kTokTile / 4)                                                 \
    SWIN_LAUNCH_CHECK();                                                                                                  \
    return 0;                                                                                                             \
  }
  NCHW_CASES(NCHW_FWD)
#undef NCHW_FWD
#undef NCHW_FWD_LAUNCH
  set_error("ln_nchw: unsupported C %d", C);
  return -EINVAL;
}
extern "C" int swin_ln_nchw_bwd(const float* dout, const float* x, const float* gamma, const float* mean, const float* rstd,
                                float* dx, float* dgamma, float* dbeta, int B, int L, int C, void* stream) {
  SWIN_REQUIRE(B > 0 && L > 0 && C > 0 && C % 4 == 0 && C <= 1024, "ln_nchw: bad shape (C %% 4 == 0, C <= 1024)");
  SWIN_REQUIRE(dout && x && gamma && mean && rstd && dx && dgamma && dbeta, "ln_nchw_bwd: null pointer");
  SWIN_REQUIRE(aligned16(x) && aligned16(dx) && aligned16(gamma), "ln_nchw_bwd: alignment");
  // Measured on the Swin-T / Swin-B stage shapes (profiles/r02/ln_nchw_bwd_sweep.txt): C <= 128 (long token axis) runs best with 32-token
  // tiles, two buffers and the next tile's dout + x rows + statistics in flight while this one is processed.  Wider rows come
  // with few tokens per image; there shared memory is better spent on residency than on staging x: 16-token dout tiles (64-byte
  // row segments), x read from global memory, two buffers -- except the 768-wide rows (128 registers, two blocks per SM),
  // where one buffer per block measured faster (2.26 vs 1.91 TB/s).
  const int stage_x = C <= 128 ? 1 : 0;
  const int tok = C <= 128 ? kTokTile : kTokTile / 2;
  const int nbuf = (C > 512 && C <= 768) ? 1 : 2;
  size_t smem = ((size_t)nbuf * ((size_t)C * (tok + 1) + (stage_x ? (size_t)tok * C + 2 * tok : 0) + 4) + 2 * C) * sizeof(float);
  SWIN_REQUIRE(smem <= 212 * 1024, "ln_nchw_bwd: tile does not fit shared memory");
  int tiles = ceil_div(L, tok);
  // tiles per block: ~8-24 so the dgamma/dbeta atomics amortise, chosen so that the grid fills the resident-block capacity
  // (<= 3 blocks per SM by registers, fewer when the tile buffers are large) in whole waves — 5.05 waves ran as 6
  int bps = (int)((227 * 1024) / (smem + 1024));
  if (bps > 3) bps = 3;
  if (bps < 1) bps = 1;
  const int cap = kNumSMs * bps;
  int tpb = 1;
  double best = -1.0;
  for (int cand = 1; cand <= 32; ++cand) {
    const long long blocks = (long long)ceil_div(tiles, cand) * B;
    double eff = (double)blocks / (double)(ceil_div64(blocks, cap) * cap);
    if (cand >= 8) eff += 0.05;                     // prefer enough tiles per block to amortise the closing atomics
    if (blocks < cap && cand > 1) break;            // small problem: keep every SM busy rather than batching tiles
    if (eff > best + 1e-9) { best = eff; tpb = cand; }
  }
  dim3 grid(ceil_div(tiles, tpb), B);
  int G, vpl;
  nchw_shape(C / 4, &G, &vpl);
#define NCHW_BWD_LAUNCH(V, GG, TK, SXV, cond)                                                                             \
  if (cond) {                                                                                                             \
    const int ar = ensure_dyn_smem((const void*)ln_nchw_bwd_kernel<V, GG, TK, SXV>, 212 * 1024);                          \
    if (ar) return ar;                                                                                                    \
    ln_nchw_bwd_kernel<V, GG, TK, SXV><<<grid, 256, smem, (cudaStream_t)stream>>>(dout, x, gamma, mean, rstd, dx, dgamma, dbeta, L, C, tpb, nbuf); \
  }
#define NCHW_BWD(V, GG)                                                                                                   \
  if (G == GG && vpl <= V) {                                                                                              \
    NCHW_BWD_LAUNCH(V, GG, kTokTile, true, stage_x)                                                                       \
    NCHW_BWD_LAUNCH(V, GG, kTokTile / 2, false, !stage_x)                                                                 \
    SWIN_LAUNCH_CHECK();                                                                                                  \
    return 0;                                                                                                             \
  }
  NCHW_CASES(NCHW_BWD)
#undef NCHW_BWD
#undef NCHW_BWD_LAUNCH
  set_error("ln_nchw: unsupported C %d", C);
  return -EINVAL;
}
