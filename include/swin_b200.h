/*
 * swin_b200 — C ABI of the B200 (sm_100a) kernels behind the Swin backbone's shifted-window
 * attention path.  This is the drop-in boundary: plain pointers and sizes, no torch types.
 *
 * The reference (an mmdetection 2.11 fork) is pure Python and ships no native interface; the
 * "FFI" for this path is the set of torch calls made by
 *     mmdet/models/backbones/swin_transformer.py            (cited below as REF:line)
 * Each entry point names the reference lines it replaces.  The Python package `swin_b200`
 * binds these with ctypes (swin_b200/_lib.py); INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function ENQUEUES work on `stream` (a cudaStream_t passed as void*) and returns:
 *       0 on success, a negative errno-style code (-EINVAL bad shape/alignment/dtype,
 *       -ENOTSUP device is not sm_100) or a positive cudaError_t.  Nothing throws, allocates
 *       device memory or synchronises.  swin_last_error() returns a thread-local message.
 *   - all device pointers must be 16-byte aligned, tensors contiguous unless a leading
 *     dimension is given.  Entry points are re-entrant (forward runs on the Python main
 *     thread, backward on autograd's worker thread).
 *   - dtypes: SWIN_F32 = 0, SWIN_BF16 = 1.  "bf16" kernels take bf16 operands, accumulate and
 *     normalise in fp32.  The residual stream and LayerNorm statistics are always fp32.
 *   - geometry: B images, H x W tokens per image (token-major (B, H*W, C)), window size ws,
 *     shift in {0, ws/2}; Hp, Wp = H, W rounded up to a multiple of ws; nW = Hp/ws * Wp/ws;
 *     N = ws*ws; window slots are (B*nW, N, C).
 */
#ifndef SWIN_B200_H_
#define SWIN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWIN_B200_VERSION 100 /* 0.1.0 */

enum { SWIN_F32 = 0, SWIN_BF16 = 1 };

int swin_version(void);
const char* swin_last_error(void);
/* 0 if `device` is compute capability 10.x, else -ENOTSUP. */
int swin_device_check(int device);
/* SMs the persistent kernels (tcgen05 GEMM, attention) leave free, e.g. for NCCL's CTAs when the gradient all-reduce of
 * mmdet/apis/train.py:91-99 overlaps backward: their grids become (SM count - n) CTAs instead of one per SM.  n < 0 only
 * queries.  Returns the previous value (initially 0).  Kernel attributes and the SM count are tracked per device. */
int swin_sm_reserve(int n);

/* ---------------------------------------------------------------- index ops (bit-exact) */

/* window_partition, REF:41-53.  x (B,Hp,Wp,C) -> win (B*nW, ws, ws, C); elem_bytes in {2,4}. */
int swin_window_partition(const void* x, void* win, int B, int Hp, int Wp, int C, int ws, int elem_bytes, void* stream);
/* window_reverse, REF:56-70.  win (B*nW, ws, ws, C) -> x (B,Hp,Wp,C). */
int swin_window_reverse(const void* win, void* x, int B, int Hp, int Wp, int C, int ws, int elem_bytes, void* stream);
/* F.pad + torch.roll(-shift) + window_partition, REF:214-231.  x (B,H*W,C) -> xw (B*nW, N, C); pad slots = 0. */
int swin_window_gather(const void* x, void* xw, int B, int H, int W, int C, int ws, int shift, int elem_bytes, void* stream);
/* window_reverse + torch.roll(+shift) + crop, REF:236-249.  xw (B*nW,N,C) -> x (B,H*W,C). */
int swin_window_scatter(const void* xw, void* x, int B, int H, int W, int C, int ws, int shift, int elem_bytes, void* stream);
/* SW-MSA mask of BasicLayer.forward, REF:370-389.  mask (nW,N,N) fp32 in {0,-100}. */
int swin_shift_mask(float* mask, int H, int W, int ws, int shift, void* stream);
/* flags[w] = 1 if mask[w] (N x N) has any non-zero entry, else 0  (lets the attention kernels skip all-zero masks). */
int swin_mask_nonzero(const float* mask, int32_t* flags, int nW, int N, void* stream);
/* relative_position_bias gather, REF:135-137: table ((2ws-1)^2, nH) -> bias (nH,N,N) fp32. */
int swin_rel_bias_expand(const float* table, float* bias, int nH, int ws, void* stream);
/* its transpose (index_put accumulate of the autograd backward): dtable += scatter(dbias). */
int swin_rel_bias_reduce(const float* dbias, float* dtable, int nH, int ws, void* stream);

/* ---------------------------------------------------------------- LayerNorm family
 * mode 0: plain rows           x (rows, C)                       -> y (rows, C)        REF:253 norm2, :620
 * mode 1: LN + pad/roll/partition  x (B,H*W,C)                   -> y (B*nW, N, C)     REF:211-231
 * mode 2: PatchMerging 2x2 gather + LN(4C)  x (B,H*W,C)          -> y (B*H2*W2, 4C)    REF:281-295
 * x is fp32; y has dtype y_dtype.  mean/rstd: fp32 per LN row (mode 0/1: B*H*W, mode 2: B*H2*W2).
 */
typedef struct swin_ln_args {
  int mode;
  int B, H, W, C;      /* mode 0: rows = B*H*W with C = row width                                  */
  int ws, shift;       /* mode 1 only                                                              */
  float eps;
  int y_dtype;         /* SWIN_F32 / SWIN_BF16 for y (fwd) and dy (bwd)                            */
  const float* x;      /* fp32 input rows                                                          */
  const float* gamma;  /* (C) or (4C)                                                              */
  const float* beta;
  void* y;             /* fwd out                                                                  */
  float* mean;         /* fwd out / bwd in                                                         */
  float* rstd;
  const void* dy;      /* bwd in, same layout as y                                                 */
  const float* dres;   /* bwd in, optional (may be NULL): gradient already flowing on x, added     */
  float* dx;           /* bwd out fp32, same layout as x                                           */
  float* dgamma;       /* bwd out, ACCUMULATED (+=) with atomics: caller zero-fills               */
  float* dbeta;
  /* bwd, modes 0 and 1, optional (NULL = off): dy2[slot(token)] = dy2_scale[b] * dx[token] in y_dtype, laid out as window
   * slots (B*nW, N, C) for (ws2, shift2) — mode 0: the drop-path-scaled, partitioned dY of the proj Linear (REF:252 backward);
   * mode 1 with ws2 = 1 (slot == token): the drop-path-scaled dY of the PREVIOUS block's fc2 Linear (REF:253 backward) —
   * emitted while dx is in registers.  Pad slots are written as zeros by the kernel (dy2 needs no initialisation).  dy2_colsum (C) += column sums. */
  void* dy2;
  const float* dy2_scale;
  float* dy2_colsum;
  int ws2, shift2;
} swin_ln_args;
int swin_ln_fwd(const swin_ln_args* a, void* stream);
int swin_ln_bwd(const swin_ln_args* a, void* stream);

/* output norm{i} + view/permute(0,3,1,2).contiguous(), REF:618-623: x (B,L,C) fp32 -> out (B,C,L) fp32 (NCHW), LN over C. */
int swin_ln_nchw_fwd(const float* x, const float* gamma, const float* beta, float* out, float* mean, float* rstd, int B, int L,
                     int C, float eps, void* stream);
/* backward: dout (B,C,L) -> dx (B,L,C); dgamma/dbeta ACCUMULATED (caller zero-fills). */
int swin_ln_nchw_bwd(const float* dout, const float* x, const float* gamma, const float* mean, const float* rstd, float* dx,
                     float* dgamma, float* dbeta, int B, int L, int C, void* stream);

/* ---------------------------------------------------------------- PatchEmbed unfold, REF:432-438
 * img (B,Cin,Hi,Wi) fp32 -> cols (B*ceil(Hi/p)*ceil(Wi/p), Cin*p*p) in `dtype`, zero-padded right/bottom, column order
 * [c][i][j] == proj.weight.view(C,-1); the conv is then swin_gemm(cols, weight).  scatter = its transpose (d img). */
int swin_patch_gather(const float* img, void* cols, int B, int Cin, int Hi, int Wi, int patch, int dtype, void* stream);
int swin_patch_scatter(const void* dcols, float* dimg, int B, int Cin, int Hi, int Wi, int patch, int dtype, void* stream);

/* ---------------------------------------------------------------- GEMM with fused epilogues
 * acc[m,n] = sum_k A[m,k] * B[n,k]
 *   a_trans = 0: A stored (M,K) row-major, leading dim lda.   a_trans = 1: A stored (K,M) row-major.
 *   b_trans = 0: B stored (N,K) row-major (an nn.Linear weight).  b_trans = 1: B stored (K,N) row-major.
 * dtype SWIN_F32: A,B fp32, FFMA path (the <=1e-4 parity mode).  SWIN_BF16: A,B bf16, tcgen05.mma
 * (TMA-fed, fp32 accumulation in TMEM).
 */
enum {
  SWIN_EPI_STORE = 0,            /* D = acc (+bias)                              qkv REF:129, reduction REF:296, dX  */
  SWIN_EPI_GELU = 1,             /* u = acc+bias ; D = gelu_erf(u) ; D2 = gelu_erf'(u)  fc1 + act, REF:33-34 (D2 saved for bwd) */
  SWIN_EPI_RESIDUAL = 2,         /* D(fp32) = aux + row_scale[b] * (acc+bias)    fc2 + drop_path + residual REF:253  */
  SWIN_EPI_SCATTER_RESIDUAL = 3, /* same, rows are window slots scattered through the REF:236-252 inverse map        */
  SWIN_EPI_DGELU = 4,            /* D = acc * aux  (aux = the saved gelu')       backward of REF:34                  */
  SWIN_EPI_ATOMIC_ADD = 5        /* D(fp32) += acc   (split-K weight gradients)                                      */
};
typedef struct swin_gemm_args {
  int dtype;            /* operand dtype: SWIN_F32 or SWIN_BF16 */
  int M, N, K;
  const void* A; int a_trans; int64_t lda;
  const void* B; int b_trans; int64_t ldb;
  int epilogue;
  const float* bias;    /* (N) fp32 or NULL */
  void* D; int d_dtype; int64_t ldd;
  void* D2;             /* GELU: second output gelu'(u), same dtype/ld as D */
  const void* aux;      /* RESIDUAL / SCATTER_RESIDUAL: fp32 residual (same layout as D); DGELU: saved gelu' (dtype d_dtype, ld ldd) */
  const float* row_scale; /* per-image drop-path multiplier (B) or NULL                                 */
  int rows_per_image;   /* RESIDUAL: H*W ; SCATTER_RESIDUAL: nW*N                                      */
  int H, W, ws, shift;  /* SCATTER_RESIDUAL geometry                                                   */
  float* colsum_a;      /* optional, ATOMIC_ADD with a_trans=1 only: colsum_a[m] += sum_k A[m,k] (fp32, caller zero-fills).
                           For dW = dY^T X this is the bias gradient sum_t dY[t,:], produced on the tensor cores by one
                           extra N=16 MMA per k-step against a constant all-ones smem tile (no extra pass over dY).     */
} swin_gemm_args;
int swin_gemm(const swin_gemm_args* a, void* stream);
/* Tile mode of the bf16 GEMM: 0 = 1-CTA tiles (128 x N) only, 1 = per-shape policy (default), 2 = CTA-pair tiles
 * (cta_group::2, 256 x N) wherever they are legal.  Results are identical in every mode; the knob exists so that tests
 * and the A/B tools can run both kernel families on the same shapes.  mode < 0 only queries.  Returns the previous mode.
 * (Initial value: environment variable SWIN_GEMM_PAIR, else 1.) */
int swin_gemm_pair_mode(int mode);
/* The tile plan swin_gemm would use for this bf16 problem (only dtype, M, N, K, a_trans, b_trans and epilogue are read; no
 * CUDA call, works without a GPU): out6 = { CTAs per tile (1 | 2), tile width N, tiles along M (of 128 x CTAs rows),
 * tiles along N, split-K factor, 64-wide k-blocks per split }. */
int swin_gemm_plan(const swin_gemm_args* a, int* out6);

/* colsum[n] += sum_m X[m,n]  (bias gradients); X (M,N) ld, dtype; colsum fp32, caller zero-fills. */
int swin_colsum(const void* X, int M, int N, int64_t ld, int dtype, float* colsum, void* stream);

/* y = row_scale[b] * x, fp32 -> dtype; mode 0 rows 1:1, mode 1 gathered into window slots (pad slots 0)
 * (the backward of the residual/scatter epilogues: dY for fc2 / proj).  If colsum != NULL it also ACCUMULATES
 * colsum[c] += sum_rows y[row, c] (the bias gradient of that Linear; fp32, caller zero-fills). */
int swin_scale_cast(const float* x, void* y, const float* row_scale, int mode, int B, int H, int W, int C, int ws,
                    int shift, int y_dtype, float* colsum, void* stream);
/* fp32 -> bf16 copy (weights shadow copies), n elements */
int swin_cast_bf16(const float* x, void* y, int64_t n, void* stream);
/* Data-parallel gradient bucket fill (replaces torch DDP's per-parameter bucket copies, mmdet/apis/train.py:91-99):
 * copies n <= SWIN_GATHER_MAX fp32 tensors src[e] (numel[e] elements) to bucket + dst_off[e] (dst_off multiples of 4
 * floats) in ONE launch.  src/dst_off/numel are HOST arrays, consumed before the call returns. */
#define SWIN_GATHER_MAX 64
int swin_grad_gather(const void* const* src, const int64_t* dst_off, const int64_t* numel, int n, float* bucket, void* stream);
/* Fused multi-tensor AdamW step (row f3): torch.optim.AdamW arithmetic — p *= 1 - lr*wd; m = lerp(m, g, 1-b1);
 * v = b2*v + (1-b2)*g*g; p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps) — for n <= SWIN_GATHER_MAX fp32 tensors
 * in ONE launch; replaces the per-tensor optimizer.step() driven by mmdet/utils/optimizer.py:22-33 with the AdamW /
 * paramwise weight-decay settings of configs/swin/mask_rcnn_swin_tiny_patch4_window7_mstrain_480-800_adamw_1x_coco.py:64-67.
 * All table arguments are HOST arrays (consumed before the call returns) of device pointers / per-tensor values.
 * w16 (array or NULL; entries may be NULL): bf16 shadow copy of the updated parameter written in the same pass.
 * Hyper-parameters are doubles so that 1 - beta rounds to fp32 exactly as torch's Python-float scalars do.
 * grad_scale multiplies every gradient first (1/world for SUM-reduced buckets, 1/loss_scale, ...). */
int swin_adamw_step(void* const* param, const void* const* grad, void* const* exp_avg, void* const* exp_avg_sq, void* const* w16,
                    const float* weight_decay, const int64_t* numel, int n, double lr, double beta1, double beta2, double eps,
                    int step, double grad_scale, void* stream);

/* ---------------------------------------------------------------- window attention core, REF:129-150
 * qkv (B_, N, 3C): columns [q|k|v] x [head] x [32].  bias (nH,N,N) fp32 (swin_rel_bias_expand).
 * mask (nW,N,N) fp32 or NULL; window w uses mask[w % nW] (REF:141-143).
 * out (B_, N, C) heads concatenated (REF:150).  lse (B_, nH, N) fp32 row log-sum-exp (saved for backward).
 * head_dim must be 32.  dtype SWIN_BF16 -> tcgen05 kernel (two 64-row padded windows per 128-row tile).
 */
typedef struct swin_attn_args {
  int dtype;
  int B_, nH, ws, nW;
  float scale;
  const void* qkv;
  const float* bias;
  const float* mask;
  const int32_t* mask_nz; /* optional (nW): 0 where mask[w] is all zeros, so the kernel skips reading it; NULL = unknown */
  int canon_nwh, canon_nww; /* optional: > 0 asserts that `mask` is the canonical SW-MSA mask of REF:370-389 for an
                               (nwh x nww) window grid with shift = ws/2; the bf16 kernel then evaluates it in closed
                               form (region ids) instead of reading the tensor.  0 = honour the tensor as given.      */
  void* out;
  float* lse;
  /* backward */
  const void* dout;   /* (B_,N,C) */
  void* dqkv;         /* (B_,N,3C) */
  float* dbias;       /* (nH,N,N) fp32, ACCUMULATED with atomics: caller zero-fills */
} swin_attn_args;
int swin_window_attn_fwd(const swin_attn_args* a, void* stream);
int swin_window_attn_bwd(const swin_attn_args* a, void* stream);

/* ---------------------------------------------------------------- fused QKV projection + window attention, REF:128-150
 * One kernel for  qkv = x Wqkv^T + bqkv ; out = softmax(scale q k^T + bias + mask) v :  the window rows x are read once and
 * Q / K / V never touch HBM (unless qkv_out is given).  bf16 operands, window 7, head_dim 32; the weight must fit in shared
 * memory next to the pipeline, or -- streamed per head through a ring of k-blocks -- leave room for the 128 x C window-pair
 * tile (swin_window_attn_qkv_supported: C <= 384, i.e. stages 0-2 of Swin-T / Swin-S, stages 0-1 of Swin-B).
 *   x     (B_*N, C)   bf16  LayerNorm'd, shifted, partitioned window rows (swin_ln_fwd mode 1)
 *   wqkv  (3C, C)     bf16  qkv.weight, rows [q|k|v] x [head] x [32] (REF:129);  bqkv (3C) fp32 or NULL
 *   bias  (nH,N,N)    fp32  (swin_rel_bias_expand);  mask / mask_nz / canon_* as in swin_attn_args
 *   out   (B_, N, C)  bf16 ; lse (B_, nH, N) fp32 or NULL (inference)
 *   qkv_out (B_, N, 3C) bf16 or NULL: also write q, k, v (what swin_window_attn_bwd reads), TMA-stored from the operand tiles
 */
typedef struct swin_attn_qkv_args {
  int B_, nH, ws, nW;
  float scale;
  const void* x;
  const void* wqkv;
  const float* bqkv;
  const float* bias;
  const float* mask;
  const int32_t* mask_nz;
  int canon_nwh, canon_nww;
  void* out;
  float* lse;
  void* qkv_out;
  void* workspace;            /* device scratch of swin_window_attn_qkv_workspace(C, nH, ws) bytes, 16-byte aligned (NULL if 0) */
  long long workspace_bytes;
} swin_attn_qkv_args;
int swin_window_attn_qkv_fwd(const swin_attn_qkv_args* a, void* stream);
/* bytes of device scratch the fused kernel needs for this shape (the padded, pre-scaled bias table of the streamed-weight mode); 0 = none */
long long swin_window_attn_qkv_workspace(int C, int nH, int ws);
/* 1 if the fused kernel can run this shape (no CUDA call), else 0. */
int swin_window_attn_qkv_supported(int C, int nH, int ws);

#ifdef __cplusplus
}
#endif
#endif /* SWIN_B200_H_ */
