"""Tensor-level wrappers over the C ABI (device pointers in, device pointers out).

PyTorch is plumbing here: it owns the buffers (caching allocator) and the stream; every
arithmetic step is a kernel from libswin_b200.so.  All wrappers raise RuntimeError on failure
and refuse non-CUDA tensors — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}
_EPI_NAMES = {L.EPI_STORE: "store", L.EPI_GELU: "gelu", L.EPI_RESIDUAL: "residual", L.EPI_SCATTER_RESIDUAL: "scatter_residual",
              L.EPI_DGELU: "dgelu", L.EPI_ATOMIC_ADD: "splitk_dW"}

# ------------------------------------------------------------------ instrumentation (bench.py)
LAUNCHES = 0          # kernels of libswin_b200.so enqueued by this process
_TIMER = None         # optional object with .begin(kind, flops, bytes) / .end()


def set_kernel_timer(timer) -> None:
    """bench.py installs a CUDA-event timer around the dominant kernel class; None disables."""
    global _TIMER
    _TIMER = timer


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _nb(*ts) -> float:
    """Bytes of the given tensors (None skipped): the algorithmic HBM traffic of a kernel is the sum over the tensors
    it must read once and write once."""
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


class _timed:
    """``with _timed(kind, flops, bytes):`` brackets one kernel launch with the installed timer (no-op otherwise)."""
    __slots__ = ("on",)

    def __init__(self, kind: str, flops: float = 0.0, nbytes: float = 0.0):
        self.on = _TIMER is not None
        if self.on:
            _TIMER.begin(kind, flops, nbytes)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.on:
            _TIMER.end()
        return False


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(*ts: Optional[torch.Tensor]) -> None:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("swin_b200 ops need CUDA tensors (no CPU fallback exists)")
        if not t.is_contiguous():
            raise RuntimeError("swin_b200 ops need contiguous tensors")
        if dev is None:
            # kernels are enqueued on the CURRENT device's current stream: the operands must live there
            dev = t.device
            if dev.index != torch.cuda.current_device():
                raise RuntimeError(f"swin_b200 ops: tensor on {dev} but the current CUDA device is {torch.cuda.current_device()} "
                                   "(wrap the call in torch.cuda.device(tensor.device))")
        elif t.device != dev:
            raise RuntimeError(f"swin_b200 ops: operands on different devices ({dev} and {t.device})")


def torch_dtype(code: int) -> torch.dtype:
    return torch.float32 if code == L.F32 else torch.bfloat16


def padded_hw(H: int, W: int, ws: int) -> Tuple[int, int]:
    return -(-H // ws) * ws, -(-W // ws) * ws


# ------------------------------------------------------------------ index ops (a1-a5)
def window_partition(x: torch.Tensor, ws: int) -> torch.Tensor:
    _chk(x)
    B, Hp, Wp, Cc = x.shape
    out = torch.empty((B * (Hp // ws) * (Wp // ws), ws, ws, Cc), dtype=x.dtype, device=x.device)
    _count()
    L.check(L.lib().swin_window_partition(_p(x), _p(out), B, Hp, Wp, Cc, ws, x.element_size(), _stream()), "window_partition")
    return out


def window_reverse(win: torch.Tensor, ws: int, Hp: int, Wp: int) -> torch.Tensor:
    _chk(win)
    Cc = win.shape[-1]
    nW = (Hp // ws) * (Wp // ws)
    B = win.shape[0] // nW
    out = torch.empty((B, Hp, Wp, Cc), dtype=win.dtype, device=win.device)
    _count()
    L.check(L.lib().swin_window_reverse(_p(win), _p(out), B, Hp, Wp, Cc, ws, win.element_size(), _stream()), "window_reverse")
    return out


def window_gather(x: torch.Tensor, H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """(B, H*W, C) -> (B*nW, ws*ws, C): pad + roll(-shift) + partition."""
    _chk(x)
    B, Lx, Cc = x.shape
    assert Lx == H * W
    Hp, Wp = padded_hw(H, W, ws)
    out = torch.empty((B * (Hp // ws) * (Wp // ws), ws * ws, Cc), dtype=x.dtype, device=x.device)
    _count()
    L.check(L.lib().swin_window_gather(_p(x), _p(out), B, H, W, Cc, ws, shift, x.element_size(), _stream()), "window_gather")
    return out


def window_scatter(xw: torch.Tensor, B: int, H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """(B*nW, ws*ws, C) -> (B, H*W, C): reverse + roll(+shift) + crop."""
    _chk(xw)
    Cc = xw.shape[-1]
    out = torch.empty((B, H * W, Cc), dtype=xw.dtype, device=xw.device)
    _count()
    L.check(L.lib().swin_window_scatter(_p(xw), _p(out), B, H, W, Cc, ws, shift, xw.element_size(), _stream()), "window_scatter")
    return out


def shift_mask(H: int, W: int, ws: int, shift: int, device) -> torch.Tensor:
    Hp, Wp = padded_hw(H, W, ws)
    nW, N = (Hp // ws) * (Wp // ws), ws * ws
    out = torch.empty((nW, N, N), dtype=torch.float32, device=device)
    _count()
    L.check(L.lib().swin_shift_mask(_p(out), H, W, ws, shift, _stream()), "shift_mask")
    return out


def mask_nonzero(mask: torch.Tensor) -> torch.Tensor:
    """(nW, N, N) fp32 -> (nW,) int32 flags: 1 where the window's mask has a non-zero entry."""
    _chk(mask)
    flags = torch.empty((mask.shape[0],), dtype=torch.int32, device=mask.device)
    _count()
    L.check(L.lib().swin_mask_nonzero(_p(mask), _p(flags), mask.shape[0], mask.shape[1], _stream()), "mask_nonzero")
    return flags


def rel_bias_expand(table: torch.Tensor, ws: int) -> torch.Tensor:
    _chk(table)
    nH = table.shape[1]
    out = torch.empty((nH, ws * ws, ws * ws), dtype=torch.float32, device=table.device)
    _count()
    L.check(L.lib().swin_rel_bias_expand(_p(table), _p(out), nH, ws, _stream()), "rel_bias_expand")
    return out


def rel_bias_reduce(dbias: torch.Tensor, ws: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out``: optional pre-zeroed ((2ws-1)^2, nH) fp32 buffer (the kernel accumulates into it)."""
    _chk(dbias, out)
    nH = dbias.shape[0]
    if out is None:
        out = torch.zeros(((2 * ws - 1) ** 2, nH), dtype=torch.float32, device=dbias.device)
    _count()
    L.check(L.lib().swin_rel_bias_reduce(_p(dbias), _p(out), nH, ws, _stream()), "rel_bias_reduce")
    return out


# ------------------------------------------------------------------ LayerNorm family
def ln_fwd(mode: int, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, B: int, H: int, W: int, Cc: int,
           ws: int, shift: int, eps: float, y_dtype: int):
    """Returns (y, mean, rstd).  mode 0 plain (rows=B*H*W), 1 LN+window gather, 2 PatchMerging gather+LN(4C)."""
    _chk(x, gamma, beta)
    assert x.dtype == torch.float32 and gamma.dtype == torch.float32
    dev = x.device
    if mode == 0:
        yshape, nstat = (B * H * W, Cc), B * H * W
    elif mode == 1:
        Hp, Wp = padded_hw(H, W, ws)
        yshape, nstat = (B * (Hp // ws) * (Wp // ws), ws * ws, Cc), B * H * W
    else:
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        yshape, nstat = (B, H2 * W2, 4 * Cc), B * H2 * W2
    y = torch.empty(yshape, dtype=torch_dtype(y_dtype), device=dev)
    mean = torch.empty((nstat,), dtype=torch.float32, device=dev)
    rstd = torch.empty((nstat,), dtype=torch.float32, device=dev)
    a = L.LnArgs(mode=mode, B=B, H=H, W=W, C=Cc, ws=ws, shift=shift, eps=eps, y_dtype=y_dtype, x=_p(x), gamma=_p(gamma),
                 beta=_p(beta), y=_p(y), mean=_p(mean), rstd=_p(rstd))
    _count()
    with _timed(f"ln_fwd mode{mode} C={Cc}", 0.0, _nb(x, y)):
        L.check(L.lib().swin_ln_fwd(C.byref(a), _stream()), "ln_fwd")
    return y, mean, rstd


def ln_bwd(mode: int, dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor,
           dres: Optional[torch.Tensor], B: int, H: int, W: int, Cc: int, ws: int, shift: int, emit_windows=None,
           dgb: Optional[torch.Tensor] = None):
    """Returns (dx fp32 like x, dgamma, dbeta).  With ``emit_windows=(ws2, shift2, row_scale)`` (modes 0 and 1) it also
    returns (dy2, colsum2): row_scale[b]*dx cast to dy's dtype and gathered into window slots, plus its column sums."""
    _chk(dy, x, gamma, mean, rstd, dres)
    dx = torch.empty_like(x)
    width = Cc * (4 if mode == 2 else 1)
    if dgb is None:                 # (3, width) pre-zeroed accumulators: dgamma, dbeta, column sums of the emitted dY
        dgb = torch.zeros((3, width), dtype=torch.float32, device=x.device)
    _chk(dgb)
    a = L.LnArgs(mode=mode, B=B, H=H, W=W, C=Cc, ws=ws, shift=shift, eps=0.0, y_dtype=_DT[dy.dtype], x=_p(x), gamma=_p(gamma),
                 mean=_p(mean), rstd=_p(rstd), dy=_p(dy), dres=_p(dres), dx=_p(dx), dgamma=_p(dgb[0]), dbeta=_p(dgb[1]))
    dy2 = None
    if emit_windows is not None:
        ws2, shift2, scale2 = emit_windows
        _chk(scale2)
        Hp, Wp = padded_hw(H, W, ws2)
        dy2 = torch.empty((B * (Hp // ws2) * (Wp // ws2) * ws2 * ws2, Cc), dtype=dy.dtype, device=x.device)   # the kernel zeroes the pad slots
        a.dy2, a.dy2_scale, a.dy2_colsum, a.ws2, a.shift2 = _p(dy2), _p(scale2), _p(dgb[2]), ws2, shift2
    _count()
    with _timed(f"ln_bwd mode{mode} C={Cc}{' +emit' if emit_windows is not None else ''}", 0.0, _nb(dy, x, dres, dx, dy2)):
        L.check(L.lib().swin_ln_bwd(C.byref(a), _stream()), "ln_bwd")
    if emit_windows is not None:
        return dx, dgb[0], dgb[1], dy2, dgb[2]
    return dx, dgb[0], dgb[1]


def ln_nchw_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, H: int, W: int, eps: float):
    """x (B, H*W, C) fp32 -> (out (B, C, H, W) fp32 contiguous, mean, rstd): output norm fused with the NCHW transpose."""
    _chk(x, gamma, beta)
    B, Lx, Cc = x.shape
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    mean = torch.empty((B * Lx,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((B * Lx,), dtype=torch.float32, device=x.device)
    _count()
    with _timed(f"ln_nchw_fwd C={Cc}", 0.0, _nb(x, out)):
        L.check(L.lib().swin_ln_nchw_fwd(_p(x), _p(gamma), _p(beta), _p(out), _p(mean), _p(rstd), B, Lx, Cc, eps, _stream()), "ln_nchw_fwd")
    return out, mean, rstd


def ln_nchw_bwd(dout: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor):
    _chk(dout, x, gamma, mean, rstd)
    B, Lx, Cc = x.shape
    dx = torch.empty_like(x)
    dgb = torch.zeros((2, Cc), dtype=torch.float32, device=x.device)
    _count()
    with _timed(f"ln_nchw_bwd C={Cc}", 0.0, _nb(dout, x, dx)):
        L.check(L.lib().swin_ln_nchw_bwd(_p(dout), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dgb[0]), _p(dgb[1]), B, Lx, Cc,
                                     _stream()), "ln_nchw_bwd")
    return dx, dgb[0], dgb[1]


def patch_gather(img: torch.Tensor, patch: int, dtype: int) -> torch.Tensor:
    """img (B,Cin,Hi,Wi) fp32 -> cols (B*Hh*Ww, Cin*p*p)."""
    _chk(img)
    B, Cin, Hi, Wi = img.shape
    Hh, Ww = -(-Hi // patch), -(-Wi // patch)
    cols = torch.empty((B * Hh * Ww, Cin * patch * patch), dtype=torch_dtype(dtype), device=img.device)
    _count()
    with _timed("patch_gather", 0.0, _nb(img, cols)):
        L.check(L.lib().swin_patch_gather(_p(img), _p(cols), B, Cin, Hi, Wi, patch, dtype, _stream()), "patch_gather")
    return cols


def patch_scatter(dcols: torch.Tensor, B: int, Cin: int, Hi: int, Wi: int, patch: int) -> torch.Tensor:
    _chk(dcols)
    dimg = torch.empty((B, Cin, Hi, Wi), dtype=torch.float32, device=dcols.device)
    _count()
    with _timed("patch_scatter", 0.0, _nb(dcols, dimg)):
        L.check(L.lib().swin_patch_scatter(_p(dcols), _p(dimg), B, Cin, Hi, Wi, patch, _DT[dcols.dtype], _stream()), "patch_scatter")
    return dimg


# ------------------------------------------------------------------ GEMM
def gemm(A: torch.Tensor, Bm: torch.Tensor, M: int, N: int, K: int, *, a_trans: bool = False, b_trans: bool = False,
         epilogue: int = L.EPI_STORE, bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         out_dtype: Optional[int] = None, out2: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
         row_scale: Optional[torch.Tensor] = None, rows_per_image: int = 0, geom=(0, 0, 0, 0),
         out_rows: Optional[int] = None, colsum_a: Optional[torch.Tensor] = None) -> torch.Tensor:
    """acc = A(M,K) . B(N,K)^T with a fused epilogue; see include/swin_b200.h."""
    _chk(A, Bm, bias, out, out2, aux, row_scale, colsum_a)
    dt = _DT[A.dtype]
    assert Bm.dtype == A.dtype
    if out is None:
        od = dt if out_dtype is None else out_dtype
        out = torch.empty((M if out_rows is None else out_rows, N), dtype=torch_dtype(od), device=A.device)
    a = L.GemmArgs(dtype=dt, M=M, N=N, K=K, A=_p(A), a_trans=int(a_trans), lda=A.stride(-2) if A.dim() >= 2 else K,
                   B=_p(Bm), b_trans=int(b_trans), ldb=Bm.stride(-2), epilogue=epilogue, bias=_p(bias), D=_p(out),
                   d_dtype=_DT[out.dtype], ldd=N, D2=_p(out2), aux=_p(aux), row_scale=_p(row_scale),
                   rows_per_image=rows_per_image, H=geom[0], W=geom[1], ws=geom[2], shift=geom[3], colsum_a=_p(colsum_a))
    _count()
    if _TIMER is not None and dt == L.BF16:
        # algorithmic bytes: both operands once, the output once (+ a second output / the aux tile where the epilogue has one);
        # the split-K weight gradient writes N x M fp32 once
        nbytes = _nb(A, Bm, out2, aux) + (_nb(out) if out_rows is None else float(M) * N * out.element_size())
        kind = f"gemm_tc {_EPI_NAMES.get(epilogue, epilogue)}{'/At' if a_trans else ''}{'/Bt' if b_trans else ''} M={M} N={N} K={K}"
        with _timed(kind, 2.0 * M * N * K, nbytes):
            L.check(L.lib().swin_gemm(C.byref(a), _stream()), "gemm")
    else:
        L.check(L.lib().swin_gemm(C.byref(a), _stream()), "gemm")
    return out


def colsum(X: torch.Tensor) -> torch.Tensor:
    _chk(X)
    X2 = X.reshape(-1, X.shape[-1])
    out = torch.zeros((X2.shape[1],), dtype=torch.float32, device=X.device)
    _count()
    with _timed("colsum", 0.0, _nb(X2)):
        L.check(L.lib().swin_colsum(_p(X2), X2.shape[0], X2.shape[1], X2.shape[1], _DT[X.dtype], _p(out), _stream()), "colsum")
    return out


def scale_cast(x: torch.Tensor, row_scale: Optional[torch.Tensor], mode: int, B: int, H: int, W: int, Cc: int, ws: int,
               shift: int, y_dtype: int, want_colsum: bool = False, colsum_out: Optional[torch.Tensor] = None):
    """fp32 -> y_dtype cast with per-image scale (mode 1: gathered into window slots).  With want_colsum also returns the
    fp32 column sums of the result (the bias gradient), computed in the same pass."""
    _chk(x, row_scale)
    assert x.dtype == torch.float32
    if mode == 1:
        Hp, Wp = padded_hw(H, W, ws)
        rows = B * (Hp // ws) * (Wp // ws) * ws * ws
    else:
        rows = B * H * W
    y = torch.empty((rows, Cc), dtype=torch_dtype(y_dtype), device=x.device)
    cs = None
    if want_colsum:               # colsum_out: optional pre-zeroed (C,) fp32 accumulator
        cs = colsum_out if colsum_out is not None else torch.zeros((Cc,), dtype=torch.float32, device=x.device)
        _chk(cs)
    _count()
    with _timed(f"scale_cast mode{mode} C={Cc}", 0.0, _nb(x, y)):
        L.check(L.lib().swin_scale_cast(_p(x), _p(y), _p(row_scale), mode, B, H, W, Cc, ws, shift, y_dtype, _p(cs), _stream()), "scale_cast")
    return (y, cs) if want_colsum else y


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _chk(x)
    assert x.dtype == torch.float32
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _count()
    L.check(L.lib().swin_cast_bf16(_p(x), _p(y), x.numel(), _stream()), "cast_bf16")
    return y


def grad_gather(tensors, offsets, bucket: torch.Tensor) -> None:
    """Copies fp32 ``tensors[i]`` (contiguous) to ``bucket[offsets[i] : offsets[i] + numel]``, GATHER_MAX tensors per launch."""
    _chk(bucket, *tensors)
    assert bucket.dtype == torch.float32
    for i0 in range(0, len(tensors), L.GATHER_MAX):
        ts, offs = tensors[i0:i0 + L.GATHER_MAX], offsets[i0:i0 + L.GATHER_MAX]
        n = len(ts)
        src = (C.c_void_p * n)(*[t.data_ptr() for t in ts])
        off = (C.c_int64 * n)(*offs)
        num = (C.c_int64 * n)(*[t.numel() for t in ts])
        _count()
        with _timed("grad_gather", 0.0, 2.0 * _nb(*ts)):
            L.check(L.lib().swin_grad_gather(src, off, num, n, _p(bucket), _stream()), "grad_gather")


def adamw_step(params, grads, exp_avgs, exp_avg_sqs, shadows, weight_decays, lr: float, beta1: float, beta2: float,
               eps: float, step: int, grad_scale: float = 1.0) -> None:
    """Fused AdamW over lists of fp32 CUDA tensors (``shadows[i]``: bf16 copy to refresh, or None), GATHER_MAX per launch."""
    _chk(*params, *grads, *exp_avgs, *exp_avg_sqs, *[s for s in shadows if s is not None])
    for i0 in range(0, len(params), L.GATHER_MAX):
        sl = slice(i0, i0 + L.GATHER_MAX)
        ps, gs, ms, vs, ws, wd = params[sl], grads[sl], exp_avgs[sl], exp_avg_sqs[sl], shadows[sl], weight_decays[sl]
        n = len(ps)
        arr = lambda ts: (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in ts])
        _count()
        with _timed("adamw_step", 0.0, 3.0 * _nb(*ps) + _nb(*gs) + 3.0 * _nb(*ps) + _nb(*[w for w in ws if w is not None])):
            L.check(L.lib().swin_adamw_step(arr(ps), arr(gs), arr(ms), arr(vs), arr(ws), (C.c_float * n)(*wd),
                                            (C.c_int64 * n)(*[t.numel() for t in ps]), n, lr, beta1, beta2, eps, step,
                                            grad_scale, _stream()), "adamw_step")


# ------------------------------------------------------------------ window attention core
def window_attn_fwd(qkv: torch.Tensor, bias: torch.Tensor, mask: Optional[torch.Tensor], B_: int, nH: int, ws: int,
                    scale: float, mask_nz: Optional[torch.Tensor] = None, canon=(0, 0)):
    _chk(qkv, bias, mask, mask_nz)
    N, Cc = ws * ws, nH * 32
    out = torch.empty((B_, N, Cc), dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty((B_, nH, N), dtype=torch.float32, device=qkv.device)
    a = L.AttnArgs(dtype=_DT[qkv.dtype], B_=B_, nH=nH, ws=ws, nW=0 if mask is None else mask.shape[0], scale=scale,
                   qkv=_p(qkv), bias=_p(bias), mask=_p(mask), mask_nz=_p(mask_nz), canon_nwh=canon[0], canon_nww=canon[1], out=_p(out), lse=_p(lse))
    _count()
    with _timed(f"attn_fwd nH={nH}", 307328.0 * B_ * nH if ws == 7 else 0.0, _nb(qkv, out)):
        L.check(L.lib().swin_window_attn_fwd(C.byref(a), _stream()), "window_attn_fwd")
    return out, lse


def window_attn_bwd(qkv: torch.Tensor, out: torch.Tensor, dout: torch.Tensor, lse: torch.Tensor, bias: torch.Tensor,
                    mask: Optional[torch.Tensor], B_: int, nH: int, ws: int, scale: float,
                    mask_nz: Optional[torch.Tensor] = None, canon=(0, 0), dbias: Optional[torch.Tensor] = None):
    """``dbias``: optional pre-zeroed (nH, N, N) fp32 accumulator."""
    _chk(qkv, out, dout, lse, bias, mask, mask_nz, dbias)
    N = ws * ws
    dqkv = torch.empty_like(qkv)
    if dbias is None:
        dbias = torch.zeros((nH, N, N), dtype=torch.float32, device=qkv.device)
    a = L.AttnArgs(dtype=_DT[qkv.dtype], B_=B_, nH=nH, ws=ws, nW=0 if mask is None else mask.shape[0], scale=scale,
                   qkv=_p(qkv), bias=_p(bias), mask=_p(mask), mask_nz=_p(mask_nz), canon_nwh=canon[0], canon_nww=canon[1], out=_p(out), lse=_p(lse), dout=_p(dout), dqkv=_p(dqkv),
                   dbias=_p(dbias))
    _count()
    with _timed(f"attn_bwd nH={nH}", 768320.0 * B_ * nH if ws == 7 else 0.0, _nb(qkv, dout, dqkv)):
        L.check(L.lib().swin_window_attn_bwd(C.byref(a), _stream()), "window_attn_bwd")
    return dqkv, dbias


def window_attn_qkv_supported(Cc: int, nH: int, ws: int) -> bool:
    return bool(L.lib().swin_window_attn_qkv_supported(Cc, nH, ws))


def window_attn_qkv_fwd(xw: torch.Tensor, wqkv: torch.Tensor, bqkv: Optional[torch.Tensor], bias: torch.Tensor,
                        mask: Optional[torch.Tensor], B_: int, nH: int, ws: int, scale: float,
                        mask_nz: Optional[torch.Tensor] = None, canon=(0, 0), want_qkv: bool = True, want_lse: bool = True):
    """Fused qkv Linear + window attention (REF:128-150): xw (B_*N, C) bf16 -> (out (B_, N, C), lse, qkv or None)."""
    _chk(xw, wqkv, bqkv, bias, mask, mask_nz)
    assert xw.dtype == torch.bfloat16 and wqkv.dtype == torch.bfloat16
    N, Cc = ws * ws, nH * 32
    out = torch.empty((B_, N, Cc), dtype=xw.dtype, device=xw.device)
    lse = torch.empty((B_, nH, N), dtype=torch.float32, device=xw.device) if want_lse else None
    qkv = torch.empty((B_, N, 3 * Cc), dtype=xw.dtype, device=xw.device) if want_qkv else None
    wsb = int(L.lib().swin_window_attn_qkv_workspace(Cc, nH, ws))
    wsp = torch.empty((wsb // 4,), dtype=torch.float32, device=xw.device) if wsb else None
    a = L.AttnQkvArgs(B_=B_, nH=nH, ws=ws, nW=0 if mask is None else mask.shape[0], scale=scale, x=_p(xw), wqkv=_p(wqkv), bqkv=_p(bqkv),
                      bias=_p(bias), mask=_p(mask), mask_nz=_p(mask_nz), canon_nwh=canon[0], canon_nww=canon[1], out=_p(out), lse=_p(lse),
                      qkv_out=_p(qkv), workspace=_p(wsp), workspace_bytes=wsb)
    _count(2 if wsb else 1)
    flops = 2.0 * B_ * N * Cc * 3 * Cc + 307328.0 * B_ * nH
    with _timed(f"attn_qkv_fwd nH={nH}", flops, _nb(xw, out, qkv)):
        L.check(L.lib().swin_window_attn_qkv_fwd(C.byref(a), _stream()), "window_attn_qkv_fwd")
    return out, lse, qkv
