"""autograd.Functions that chain the sm_100a kernels into the reference's operators.

Each Function is one reference operator (forward + hand-written backward):
  SwinBlockFn        SwinTransformerBlock.forward   mmdet/models/backbones/swin_transformer.py:198-255
  WindowAttentionFn  WindowAttention.forward        :121-153
  MlpFn              Mlp.forward                    :32-38 (stand-alone calls; inside a block the MLP is part of SwinBlockFn)
  PatchMergingFn     PatchMerging.forward           :271-298
Gradients of parameters are RETURNED (not accumulated in place) so AccumulateGrad hooks — and
therefore the bucketed NCCL all-reduce in ddp.py — fire per parameter during backward.

Precision modes (``dt``): F32 = everything fp32 on the FFMA kernels (<=1e-4 parity mode);
BF16 = bf16 GEMM/attention operands on tcgen05 with fp32 accumulation, fp32 residual stream,
fp32 LayerNorm / softmax statistics and fp32 master weights (the reference's AMP semantics).
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch

from . import _lib as L
from . import ops

_W16 = {}     # id(param) -> (weakref(param), version, bf16 copy)
# inference: qkv Linear + window attention as ONE kernel (SWIN_FUSE_QKV=0: the two-kernel path, for A/B runs).  The kernel runs
# every C <= 384; it is the default only where it is measured faster than the GEMM + attention pair (resident weights, C <= 128:
# profiles/r02/attn_qkv.txt) -- SWIN_FUSE_QKV=2 forces it wherever it is supported.
_FUSE_ENV = __import__("os").environ.get("SWIN_FUSE_QKV", "1")
FUSE_QKV_ATTENTION = _FUSE_ENV != "0"
FUSE_QKV_MAX_C = 384 if _FUSE_ENV == "2" else 128


def _fuse_qkv(Cc: int, nH: int, ws: int) -> bool:
    return FUSE_QKV_ATTENTION and Cc <= FUSE_QKV_MAX_C and ops.window_attn_qkv_supported(Cc, nH, ws)


# Backward: the four weight-gradient GEMMs of a block (split-K, atomics into pre-zeroed buffers) feed nothing downstream in the
# block, so they can run on a side stream and fill the tail of the dX chain's persistent kernels (SWIN_DW_STREAM=1; joined before
# backward returns, so autograd / the DDP hooks see ordinary main-stream tensors; under CUDA-graph capture the fork / join become
# graph edges).
DW_SIDE_STREAM = __import__("os").environ.get("SWIN_DW_STREAM", "0") == "1"
_SIDE_STREAMS = {}


class _SideLane:
    """Runs callables on a per-device side stream, each after everything enqueued so far on the current stream."""

    def __init__(self, device):
        self.enabled = DW_SIDE_STREAM and device.type == "cuda"
        if not self.enabled:
            return
        self.main = torch.cuda.current_stream(device)
        key = (device.index, self.main.cuda_stream)
        if key not in _SIDE_STREAMS:
            _SIDE_STREAMS[key] = torch.cuda.Stream(device)
        self.side = _SIDE_STREAMS[key]
        self.used = False

    def run(self, fn):
        if not self.enabled:
            return fn()
        ev = torch.cuda.Event()
        ev.record(self.main)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            fn()
        self.used = True

    def join(self):
        if self.enabled and self.used:
            ev = torch.cuda.Event()
            ev.record(self.side)
            self.main.wait_event(ev)


def _w(param: torch.Tensor, dt: int) -> torch.Tensor:
    """Operand copy of a weight in the compute dtype.  bf16 shadow copies are cached per parameter OBJECT (weakly
    referenced, so an address reused by another model's parameter can never alias) and re-cast when the
    parameter's version counter changes (optimizer step, load_state_dict)."""
    if dt == L.F32:
        return param.detach()
    key = id(param)
    hit = _W16.get(key)
    ver = param._version
    if hit is not None and hit[0]() is param and hit[1] == ver and hit[2].device == param.device:
        return hit[2]
    w16 = ops.cast_bf16(param.detach().contiguous())
    _W16[key] = (weakref.ref(param, lambda _r, k=key: _W16.pop(k, None)), ver, w16)
    return w16


def _zeros_flat(device, *shapes):
    """Several zero-initialised fp32 tensors carved (16-byte aligned) out of ONE allocation / ONE memset."""
    sizes = [(int(torch.Size(sh).numel()) + 3) // 4 * 4 for sh in shapes]
    flat = torch.zeros((sum(sizes),), dtype=torch.float32, device=device)
    out, off = [], 0
    for sh, n in zip(shapes, sizes):
        out.append(flat[off:off + int(torch.Size(sh).numel())].view(sh))
        off += n
    return out


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class BlockLink:
    """Hand-over between two consecutive blocks of a stage in backward.  Block b+1's last backward kernel (LN1 backward, which
    writes dx = the gradient of block b's output) can emit, from registers, what block b's backward would otherwise produce
    with a separate pass over that gradient: the drop-path-scaled, compute-dtype copy (dY of fc2) and its column sums
    (d fc2.bias).  ``s2`` is set by block b in forward; ``dy2 / colsum / key`` are deposited by block b+1 in backward and
    consumed (then cleared) by block b only if its incoming gradient is that very tensor, unmodified."""
    __slots__ = ("s2", "dy2", "colsum", "key")

    def __init__(self):
        self.s2 = None
        self.dy2 = self.colsum = self.key = None

    def deposit(self, dx, dy2, colsum):
        self.dy2, self.colsum, self.key = dy2, colsum, (dx.data_ptr(), tuple(dx.shape), dx._version)

    def take(self, dx2):
        dy2, colsum, key = self.dy2, self.colsum, self.key
        self.dy2 = self.colsum = self.key = None
        if dy2 is not None and key == (dx2.data_ptr(), tuple(dx2.shape), dx2._version):
            return dy2, colsum
        return None


class SwinBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n1w, n1b, table, qkvw, qkvb, projw, projb, n2w, n2b, fc1w, fc1b, fc2w, fc2b, mask, mask_nz, s1, s2,
                H, W, ws, shift, nH, scale, dt, eps, canon=(0, 0), recv=None, send=None, infer=False):
        """``infer``: the caller ran under torch.no_grad() (grad mode is always off INSIDE Function.forward, so the module passes
        it in): nothing is needed for a backward, which lets the attention branch use the fused QKV + attention kernel."""
        B, Lx, Cc = x.shape
        x = _f32c(x)
        geom = (H, W, ws, shift)
        hid = fc1w.shape[0]
        T = B * Lx
        # attention branch: LN1 + pad/roll/partition -> qkv -> window attention -> proj + reverse/roll/crop + residual
        xw, mean1, rstd1 = ops.ln_fwd(1, x, n1w.detach(), n1b.detach(), B, H, W, Cc, ws, shift, eps, dt)
        Tp = xw.shape[0] * xw.shape[1]
        bias = ops.rel_bias_expand(table.detach().contiguous(), ws)
        if dt == L.BF16 and infer and _fuse_qkv(Cc, nH, ws):
            # inference (no backward will run): qkv projection inside the attention kernel -- the window rows are read once and
            # q, k, v never reach HBM.  With a backward to feed, q / k / v must be written anyway and the two-kernel chain is
            # as fast (measured, profiles/r02/attn_qkv.txt), so training keeps it.
            o, lse, qkv = ops.window_attn_qkv_fwd(xw.view(Tp, Cc), _w(qkvw, dt), None if qkvb is None else qkvb.detach(), bias, mask,
                                                  Tp // (ws * ws), nH, ws, scale, mask_nz, canon, want_qkv=False, want_lse=False)
        else:
            qkv = ops.gemm(xw, _w(qkvw, dt), Tp, 3 * Cc, Cc, bias=None if qkvb is None else qkvb.detach())
            o, lse = ops.window_attn_fwd(qkv.view(-1, ws * ws, 3 * Cc), bias, mask, Tp // (ws * ws), nH, ws, scale, mask_nz, canon)
        x1 = torch.empty_like(x)
        ops.gemm(o, _w(projw, dt), Tp, Cc, Cc, bias=projb.detach(), epilogue=L.EPI_SCATTER_RESIDUAL, out=x1, aux=x,
                 row_scale=s1, geom=geom)
        # MLP branch: LN2 -> fc1 + GELU -> fc2 + residual
        xn, mean2, rstd2 = ops.ln_fwd(0, x1, n2w.detach(), n2b.detach(), B, H, W, Cc, 1, 0, eps, dt)
        u = torch.empty((T, hid), dtype=xn.dtype, device=x.device)      # gelu'(fc1 pre-activation), saved for backward
        h = ops.gemm(xn, _w(fc1w, dt), T, hid, Cc, bias=fc1b.detach(), epilogue=L.EPI_GELU, out2=u)
        x2 = torch.empty_like(x)
        ops.gemm(h, _w(fc2w, dt), T, Cc, hid, bias=fc2b.detach(), epilogue=L.EPI_RESIDUAL, out=x2, aux=x1, row_scale=s2,
                 rows_per_image=Lx)
        ctx.save_for_backward(x, n1w, table, qkvw, projw, n2w, fc1w, fc2w, mask, mask_nz, s1, s2,
                              xw, mean1, rstd1, qkv, bias, o, lse, x1, xn, mean2, rstd2, u, h)
        ctx.cfg = (B, H, W, Cc, ws, shift, nH, scale, dt, hid, qkvb is not None, canon)
        # recv: link to the NEXT block (it deposits this block's fc2 dY); send: link to the PREVIOUS block (this block deposits)
        ctx.recv, ctx.send = recv, send
        if recv is not None:
            recv.s2 = s2
        return x2

    @staticmethod
    def backward(ctx, dx2):
        (x, n1w, table, qkvw, projw, n2w, fc1w, fc2w, mask, mask_nz, s1, s2,
         xw, mean1, rstd1, qkv, bias, o, lse, x1, xn, mean2, rstd2, u, h) = ctx.saved_tensors
        B, H, W, Cc, ws, shift, nH, scale, dt, hid, has_qkvb, canon = ctx.cfg
        T = B * H * W
        N = ws * ws
        Tp = xw.shape[0] * N
        dx2 = _f32c(dx2)
        # ---- MLP branch
        # every accumulator of this block's backward (weight / bias / LayerNorm / bias-table gradients) comes out of ONE
        # zero-filled allocation: one memset per block instead of seven
        dfc2w, dfc1w, dprojw, dqkvw, dfc1b, dqkvb_buf, dfc2b_buf, dgb2, dgb1, dbias_buf, dtable_buf = _zeros_flat(
            dx2.device, tuple(fc2w.shape), tuple(fc1w.shape), tuple(projw.shape), tuple(qkvw.shape), (hid,), (3 * Cc,),
            (Cc,), (3, Cc), (3, Cc), (nH, N, N), tuple(table.shape))
        lane = _SideLane(dx2.device)
        got = ctx.recv.take(dx2) if ctx.recv is not None else None
        if got is not None:
            dy2, dfc2b = got                         # emitted by the next block's LN1 backward (BlockLink)
        else:
            dy2, dfc2b = ops.scale_cast(dx2, s2, 0, B, H, W, Cc, 1, 0, dt, want_colsum=True, colsum_out=dfc2b_buf)  # (T, C) + bias grad
        # weight gradients of frozen Linears (requires_grad False: frozen_stages, REF:557-572) are not computed at all
        need = ctx.needs_input_grad
        n_qkv, n_proj, n_fc1, n_fc2 = need[4] or need[5], need[6] or need[7], need[10] or need[11], need[12] or need[13]
        if n_fc2:
            lane.run(lambda: ops.gemm(dy2, h, Cc, hid, T, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dfc2w))
        du = ops.gemm(dy2, _w(fc2w, dt), T, hid, Cc, b_trans=True, epilogue=L.EPI_DGELU, aux=u)
        if n_fc1:
            lane.run(lambda: ops.gemm(du, xn, hid, Cc, T, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dfc1w, colsum_a=dfc1b))
        dxn = ops.gemm(du, _w(fc1w, dt), T, Cc, hid, b_trans=True)
        # LN2 backward + residual-gradient add; the same kernel also emits dY of the proj Linear (drop-path scaled,
        # cast and partitioned into window slots) and its column sums (= d proj.bias)
        dx1, dn2w, dn2b, dy1, dprojb = ops.ln_bwd(0, dxn, x1, n2w.detach(), mean2, rstd2, dx2, B, H, W, Cc, 1, 0,
                                                  emit_windows=(ws, shift, s1), dgb=dgb2)
        # ---- attention branch
        if n_proj:
            lane.run(lambda: ops.gemm(dy1, o, Cc, Cc, Tp, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dprojw))
        do = ops.gemm(dy1, _w(projw, dt), Tp, Cc, Cc, b_trans=True)
        dqkv, dbias = ops.window_attn_bwd(qkv.view(-1, N, 3 * Cc), o, do.view(-1, N, Cc), lse, bias, mask, Tp // N, nH, ws, scale, mask_nz, canon,
                                          dbias=dbias_buf)
        dtable = ops.rel_bias_reduce(dbias, ws, out=dtable_buf) if need[3] else None
        dqkvb = dqkvb_buf if has_qkvb else None
        if n_qkv:
            lane.run(lambda: ops.gemm(dqkv, xw, 3 * Cc, Cc, Tp, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dqkvw, colsum_a=dqkvb))
        if not (need[0] or need[1] or need[2]):
            lane.join()
            # first trainable block behind frozen stages: nothing upstream wants dx, norm1 is frozen too
            return (None, None, None, dtable, dqkvw if need[4] else None, dqkvb if need[5] else None, dprojw if need[6] else None,
                    dprojb if need[7] else None, dn2w, dn2b, dfc1w if need[10] else None, dfc1b if need[11] else None,
                    dfc2w if need[12] else None, dfc2b if need[13] else None) + (None,) * 16
        dxw = ops.gemm(dqkv.view(Tp, 3 * Cc), _w(qkvw, dt), Tp, Cc, 3 * Cc, b_trans=True)
        if ctx.send is not None:
            # also emit the previous block's fc2 dY (its drop-path scale, compute dtype, token order: "windows" of one token)
            dx, dn1w, dn1b, dyp, csp = ops.ln_bwd(1, dxw, x, n1w.detach(), mean1, rstd1, dx1, B, H, W, Cc, ws, shift,
                                                  emit_windows=(1, 0, ctx.send.s2), dgb=dgb1)
            ctx.send.deposit(dx, dyp, csp)
        else:
            dx, dn1w, dn1b = ops.ln_bwd(1, dxw, x, n1w.detach(), mean1, rstd1, dx1, B, H, W, Cc, ws, shift, dgb=dgb1)
        lane.join()
        return (dx, dn1w, dn1b, dtable, dqkvw if need[4] else None, dqkvb if need[5] else None, dprojw if need[6] else None,
                dprojb if need[7] else None, dn2w, dn2b, dfc1w if need[10] else None, dfc1b if need[11] else None,
                dfc2w if need[12] else None, dfc2b if need[13] else None) + (None,) * 16


class MlpFn(torch.autograd.Function):
    """Mlp.forward (REF:32-38) called on its own: x (..., Cin) -> fc2(GELU_erf(fc1(x))), the same two GEMM kernels the fused
    block uses (fc1 with the GELU epilogue, which also saves gelu'(u) for backward) without LayerNorm / residual."""

    @staticmethod
    def forward(ctx, x, fc1w, fc1b, fc2w, fc2b, dt):
        Cin, hid, Cout = fc1w.shape[1], fc1w.shape[0], fc2w.shape[0]
        xin = _f32c(x).view(-1, Cin)
        rows = xin.shape[0]
        xa = xin if dt == L.F32 else ops.scale_cast(xin, None, 0, 1, rows, 1, Cin, 1, 0, dt)
        xa = xa.view(rows, Cin)
        u = torch.empty((rows, hid), dtype=xa.dtype, device=x.device)
        h = ops.gemm(xa, _w(fc1w, dt), rows, hid, Cin, bias=None if fc1b is None else fc1b.detach(), epilogue=L.EPI_GELU, out2=u)
        y = ops.gemm(h, _w(fc2w, dt), rows, Cout, hid, bias=None if fc2b is None else fc2b.detach(), out_dtype=L.F32)
        ctx.save_for_backward(fc1w, fc2w, xa, u, h)
        ctx.cfg = (tuple(x.shape), x.dtype, dt, fc1b is not None, fc2b is not None)
        return y.view(tuple(x.shape[:-1]) + (Cout,)).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        fc1w, fc2w, xa, u, h = ctx.saved_tensors
        xshape, xdtype, dt, has_b1, has_b2 = ctx.cfg
        Cin, hid, Cout = fc1w.shape[1], fc1w.shape[0], fc2w.shape[0]
        rows = xa.shape[0]
        dyf = _f32c(dy).view(rows, Cout)
        dy2 = dyf if dt == L.F32 else ops.scale_cast(dyf, None, 0, 1, rows, 1, Cout, 1, 0, dt)
        dy2 = dy2.view(rows, Cout)
        dfc2b = ops.colsum(dy2) if has_b2 else None
        dfc2w = torch.zeros_like(fc2w, dtype=torch.float32)
        ops.gemm(dy2, h, Cout, hid, rows, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dfc2w)
        du = ops.gemm(dy2, _w(fc2w, dt), rows, hid, Cout, b_trans=True, epilogue=L.EPI_DGELU, aux=u)
        dfc1b = ops.colsum(du) if has_b1 else None
        dfc1w = torch.zeros_like(fc1w, dtype=torch.float32)
        ops.gemm(du, xa, hid, Cin, rows, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dfc1w)
        dx = ops.gemm(du, _w(fc1w, dt), rows, Cin, hid, b_trans=True, out_dtype=L.F32)
        return dx.view(xshape).to(xdtype), dfc1w, dfc1b, dfc2w, dfc2b, None


class WindowAttentionFn(torch.autograd.Function):
    """x_windows (B_, N, C) -> (B_, N, C): qkv Linear, attention core, proj Linear."""

    @staticmethod
    def forward(ctx, xwin, table, qkvw, qkvb, projw, projb, mask, mask_nz, canon, ws, nH, scale, dt, infer=False):
        B_, N, Cc = xwin.shape
        rows = B_ * N
        xin = _f32c(xwin)
        xw = xin if dt == L.F32 else ops.scale_cast(xin, None, 0, 1, rows, 1, Cc, 1, 0, dt)
        xw = xw.view(rows, Cc)
        bias = ops.rel_bias_expand(table.detach().contiguous(), ws)
        if dt == L.BF16 and infer and _fuse_qkv(Cc, nH, ws):
            o, lse, qkv = ops.window_attn_qkv_fwd(xw, _w(qkvw, dt), None if qkvb is None else qkvb.detach(), bias, mask, B_, nH, ws, scale,
                                                  mask_nz, canon, want_qkv=False, want_lse=False)
        else:
            qkv = ops.gemm(xw, _w(qkvw, dt), rows, 3 * Cc, Cc, bias=None if qkvb is None else qkvb.detach())
            o, lse = ops.window_attn_fwd(qkv.view(B_, N, 3 * Cc), bias, mask, B_, nH, ws, scale, mask_nz, canon)
        y = ops.gemm(o.view(rows, Cc), _w(projw, dt), rows, Cc, Cc, bias=projb.detach(), out_dtype=L.F32)
        ctx.save_for_backward(table, qkvw, projw, mask, mask_nz, xw, qkv, bias, o, lse)
        ctx.cfg = (B_, N, Cc, ws, nH, scale, dt, qkvb is not None, xwin.dtype, canon)
        return y.view(B_, N, Cc).to(xwin.dtype)

    @staticmethod
    def backward(ctx, dy):
        table, qkvw, projw, mask, mask_nz, xw, qkv, bias, o, lse = ctx.saved_tensors
        B_, N, Cc, ws, nH, scale, dt, has_qkvb, in_dtype, canon = ctx.cfg
        rows = B_ * N
        dyf = _f32c(dy)
        dy1 = dyf.view(rows, Cc) if dt == L.F32 else ops.scale_cast(dyf, None, 0, 1, rows, 1, Cc, 1, 0, dt)
        dprojb = ops.colsum(dy1)
        dprojw = torch.zeros_like(projw, dtype=torch.float32)
        ops.gemm(dy1, o.view(rows, Cc), Cc, Cc, rows, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dprojw)
        do = ops.gemm(dy1, _w(projw, dt), rows, Cc, Cc, b_trans=True)
        dqkv, dbias = ops.window_attn_bwd(qkv.view(B_, N, 3 * Cc), o, do.view(B_, N, Cc), lse, bias, mask, B_, nH, ws, scale, mask_nz, canon)
        dtable = ops.rel_bias_reduce(dbias, ws)
        dqkvb = ops.colsum(dqkv) if has_qkvb else None
        dqkvw = torch.zeros_like(qkvw, dtype=torch.float32)
        ops.gemm(dqkv.view(rows, 3 * Cc), xw, 3 * Cc, Cc, rows, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dqkvw)
        dx = ops.gemm(dqkv.view(rows, 3 * Cc), _w(qkvw, dt), rows, Cc, 3 * Cc, b_trans=True, out_dtype=L.F32)
        return dx.view(B_, N, Cc).to(in_dtype), dtable, dqkvw, dqkvb, dprojw, dprojb, None, None, None, None, None, None, None, None


class PatchMergingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, nw, nb, redw, H, W, dt, eps):
        B, Lx, Cc = x.shape
        x = _f32c(x)
        g, mean, rstd = ops.ln_fwd(2, x, nw.detach(), nb.detach(), B, H, W, Cc, 1, 0, eps, dt)
        T2 = g.shape[0] * g.shape[1]
        y = ops.gemm(g.view(T2, 4 * Cc), _w(redw, dt), T2, 2 * Cc, 4 * Cc, out_dtype=L.F32)
        ctx.save_for_backward(x, nw, redw, g, mean, rstd)
        ctx.cfg = (B, H, W, Cc, dt, T2)
        return y.view(B, g.shape[1], 2 * Cc)

    @staticmethod
    def backward(ctx, dy):
        x, nw, redw, g, mean, rstd = ctx.saved_tensors
        B, H, W, Cc, dt, T2 = ctx.cfg
        dyf = _f32c(dy).view(T2, 2 * Cc)
        dy1 = dyf if dt == L.F32 else ops.scale_cast(dyf, None, 0, 1, T2, 1, 2 * Cc, 1, 0, dt)
        dredw = torch.zeros_like(redw, dtype=torch.float32)
        ops.gemm(dy1, g.view(T2, 4 * Cc), 2 * Cc, 4 * Cc, T2, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dredw)
        dg = ops.gemm(dy1, _w(redw, dt), T2, 4 * Cc, 2 * Cc, b_trans=True)
        dx, dnw, dnb = ops.ln_bwd(2, dg, x, nw.detach(), mean, rstd, None, B, H, W, Cc, 1, 0)
        return dx, dnw, dnb, dredw, None, None, None, None


class PatchEmbedFn(torch.autograd.Function):
    """img (B,Cin,Hi,Wi) -> tokens (B, Hh*Ww, C): patch unfold + GEMM (+ LayerNorm), REF:429-445, emitted directly in
    the token-major layout the block stack consumes (no NCHW round trip)."""

    @staticmethod
    def forward(ctx, img, projw, projb, nw, nb, patch, dt, eps):
        B, Cin, Hi, Wi = img.shape
        img32 = _f32c(img)
        Cc = projw.shape[0]
        K = Cin * patch * patch
        cols = ops.patch_gather(img32, patch, dt)
        T = cols.shape[0]
        y0 = ops.gemm(cols, _w(projw, dt).view(Cc, K), T, Cc, K, bias=projb.detach(), out_dtype=L.F32)
        if nw is not None:
            y, mean, rstd = ops.ln_fwd(0, y0, nw.detach(), nb.detach(), 1, T, 1, Cc, 1, 0, eps, L.F32)
        else:
            y, mean, rstd = y0, None, None
        ctx.save_for_backward(projw, nw, cols, y0, mean, rstd)
        ctx.cfg = (B, Cin, Hi, Wi, patch, dt, Cc, K, T, ctx.needs_input_grad[0])
        return y.view(B, T // B, Cc)

    @staticmethod
    def backward(ctx, dy):
        projw, nw, cols, y0, mean, rstd = ctx.saved_tensors
        B, Cin, Hi, Wi, patch, dt, Cc, K, T, need_dimg = ctx.cfg
        dyf = _f32c(dy).view(T, Cc)
        dnw = dnb = None
        if nw is not None:
            dyf, dnw, dnb = ops.ln_bwd(0, dyf, y0, nw.detach(), mean, rstd, None, 1, T, 1, Cc, 1, 0)
        d16 = dyf if dt == L.F32 else ops.scale_cast(dyf, None, 0, 1, T, 1, Cc, 1, 0, dt)
        db = ops.colsum(d16)
        dw = torch.zeros((Cc, K), dtype=torch.float32, device=dy.device)
        ops.gemm(d16, cols, Cc, K, T, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dw)
        dimg = None
        if need_dimg:
            dcols = ops.gemm(d16, _w(projw, dt).view(Cc, K), T, K, Cc, b_trans=True, out_dtype=L.F32)
            dimg = ops.patch_scatter(dcols, B, Cin, Hi, Wi, patch)
        return dimg, dw.view_as(projw), db, dnw, dnb, None, None, None


class OutNormFn(torch.autograd.Function):
    """norm{i}(x).view(B,H,W,C).permute(0,3,1,2).contiguous()  (REF:618-623) as one kernel each way."""

    @staticmethod
    def forward(ctx, x, nw, nb, H, W, eps):
        x = _f32c(x)
        out, mean, rstd = ops.ln_nchw_fwd(x, nw.detach(), nb.detach(), H, W, eps)
        ctx.save_for_backward(x, nw, mean, rstd)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, nw, mean, rstd = ctx.saved_tensors
        dx, dg, db = ops.ln_nchw_bwd(_f32c(dout), x, nw.detach(), mean, rstd)
        return dx, dg, db, None, None, None


class OutNormMergeFn(torch.autograd.Function):
    """The two consumers of a stage output -- norm{i} + NCHW (REF:618-623) and the stage's PatchMerging (REF:258-298) -- as ONE
    autograd node.  Separately they hand autograd two full-size dx tensors that it sums with an ATen add (3 passes over
    (B, H*W, C) fp32); here the out-norm backward runs first and the PatchMerging LayerNorm backward takes its dx as the
    residual gradient (``dres``) it already knows how to add, so x's gradient is written once."""

    @staticmethod
    def forward(ctx, x, onw, onb, mnw, mnb, redw, H, W, dt, eps_out, eps_merge, recv=None):
        """``recv``: BlockLink to the FIRST block of the next stage, whose LN1 backward holds this node's incoming dy in registers
        and emits its compute-dtype copy (what ``scale_cast`` would produce here with a pass of its own)."""
        B, Lx, Cc = x.shape
        x = _f32c(x)
        out, omean, orstd = ops.ln_nchw_fwd(x, onw.detach(), onb.detach(), H, W, eps_out)
        g, mean, rstd = ops.ln_fwd(2, x, mnw.detach(), mnb.detach(), B, H, W, Cc, 1, 0, eps_merge, dt)
        T2 = g.shape[0] * g.shape[1]
        y = ops.gemm(g.view(T2, 4 * Cc), _w(redw, dt), T2, 2 * Cc, 4 * Cc, out_dtype=L.F32)
        ctx.save_for_backward(x, onw, omean, orstd, mnw, redw, g, mean, rstd)
        ctx.cfg = (B, H, W, Cc, dt, T2)
        ctx.recv = recv
        return out, y.view(B, g.shape[1], 2 * Cc)

    @staticmethod
    def backward(ctx, dout, dy):
        x, onw, omean, orstd, mnw, redw, g, mean, rstd = ctx.saved_tensors
        B, H, W, Cc, dt, T2 = ctx.cfg
        dx_out, dog, dob = ops.ln_nchw_bwd(_f32c(dout), x, onw.detach(), omean, orstd)
        dyc = _f32c(dy)
        dyf = dyc.view(T2, 2 * Cc)
        got = ctx.recv.take(dyc) if ctx.recv is not None else None
        if dt == L.F32:
            dy1 = dyf
        elif got is not None:
            dy1 = got[0].view(T2, 2 * Cc)          # emitted by the next stage's first LN1 backward (BlockLink)
        else:
            dy1 = ops.scale_cast(dyf, None, 0, 1, T2, 1, 2 * Cc, 1, 0, dt)
        dredw = torch.zeros_like(redw, dtype=torch.float32)
        ops.gemm(dy1, g.view(T2, 4 * Cc), 2 * Cc, 4 * Cc, T2, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dredw)
        dg = ops.gemm(dy1, _w(redw, dt), T2, 4 * Cc, 2 * Cc, b_trans=True)
        dx, dnw, dnb = ops.ln_bwd(2, dg, x, mnw.detach(), mean, rstd, dx_out, B, H, W, Cc, 1, 0)
        return dx, dog, dob, dnw, dnb, dredw, None, None, None, None, None, None

