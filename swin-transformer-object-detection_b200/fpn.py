"""Row f2 of the hot-path table: FPN lateral 1x1 convolutions fed directly from the backbone's token-major stage outputs.

The reference computes, per output stage, ``norm{i}(x_out).view(B,H,W,C).permute(0,3,1,2).contiguous()``
(mmdet/models/backbones/swin_transformer.py:618-623) and then, in the neck, ``lateral_conv(inputs[i])`` — a 1x1
``Conv2d(C_i, 256)`` with bias and no norm/activation (mmdet/models/necks/fpn.py:120-127, applied at :169-173;
configs/_base_/models/mask_rcnn_swin_fpn.py:21-25).  A 1x1 convolution on NCHW is a GEMM on the (B*H*W, C) token
matrix, so the pair runs here as LayerNorm (bf16 operand written once) + one tcgen05 GEMM, and the NCHW transpose of
the backbone output — forward and backward, 1.1 ms of the benchmark step — disappears.  The result is returned as a
``(B, 256, H, W)`` tensor in channels-last memory format (the GEMM's (B*H*W, 256) output viewed in place), which the
top-down additions, ``F.interpolate`` and the 3x3 ``fpn_convs`` of the neck consume unchanged.

No new parameters: ``SwinFPNLaterals`` uses the backbone's own ``norm{i}`` and the neck's own ``lateral_convs[i].conv``
modules, so checkpoints and optimizers see exactly the reference's parameter set.  CUDA only.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .functional import _f32c, _w


class NormLateralFn(torch.autograd.Function):
    """tokens (B, H*W, C) -> LayerNorm(C) -> x W^T + b  ->  (B, Cout, H, W) channels-last view."""

    @staticmethod
    def forward(ctx, x, nw, nb, lw, lb, H, W, dt, eps):
        B, Lx, Cc = x.shape
        assert Lx == H * W
        x = _f32c(x)
        T, Cout = B * Lx, lw.shape[0]
        xn, mean, rstd = ops.ln_fwd(0, x, nw.detach(), nb.detach(), B, H, W, Cc, 1, 0, eps, dt)
        w2 = _w(lw, dt).view(Cout, Cc)
        y = ops.gemm(xn.view(T, Cc), w2, T, Cout, Cc, bias=None if lb is None else lb.detach(), out_dtype=L.F32)
        ctx.save_for_backward(x, nw, lw, xn, mean, rstd)
        ctx.cfg = (B, H, W, Cc, Cout, dt, lb is not None)
        return y.view(B, H, W, Cout).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        x, nw, lw, xn, mean, rstd = ctx.saved_tensors
        B, H, W, Cc, Cout, dt, has_bias = ctx.cfg
        T = B * H * W
        dyt = _f32c(dy.permute(0, 2, 3, 1)).view(T, Cout)          # no copy when dy is channels-last
        if dt == L.F32:
            d16, dlb = dyt, (ops.colsum(dyt) if has_bias else None)
        else:
            d16, dlb = ops.scale_cast(dyt, None, 0, 1, T, 1, Cout, 1, 0, dt, want_colsum=True)
            if not has_bias:
                dlb = None
        dlw = torch.zeros((Cout, Cc), dtype=torch.float32, device=x.device)
        ops.gemm(d16, xn.view(T, Cc), Cout, Cc, T, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dlw)
        dxn = ops.gemm(d16, _w(lw, dt).view(Cout, Cc), T, Cc, Cout, b_trans=True)
        dx, dnw, dnb = ops.ln_bwd(0, dxn, x, nw.detach(), mean, rstd, None, B, H, W, Cc, 1, 0)
        return dx, dnw, dnb, dlw.view_as(lw), dlb, None, None, None, None


class SwinFPNLaterals(nn.Module):
    """``laterals = SwinFPNLaterals(backbone, fpn.lateral_convs)(img)`` == ``[l(x) for l, x in zip(lateral_convs,
    backbone(img))]`` of fpn.py:169-173, in channels-last memory format.  ``lateral_convs`` may hold mmcv ``ConvModule``s
    (their ``.conv`` is used; a lateral with norm or activation is refused) or plain 1x1 ``nn.Conv2d``s."""

    def __init__(self, backbone: nn.Module, lateral_convs: Sequence[nn.Module]):
        super().__init__()
        self.backbone = backbone
        convs: List[nn.Conv2d] = []
        for m in lateral_convs:
            if getattr(m, "norm_name", None) or getattr(m, "with_norm", False) or getattr(m, "with_activation", False):
                raise NotImplementedError("fused laterals support the reference's plain 1x1 conv + bias (no norm / activation)")
            conv = getattr(m, "conv", m)
            if not isinstance(conv, nn.Conv2d) or conv.kernel_size != (1, 1) or conv.stride != (1, 1) or conv.groups != 1:
                raise NotImplementedError("lateral convolutions must be 1x1, stride 1, groups 1")
            convs.append(conv)
        if len(convs) != len(backbone.out_indices):
            raise ValueError("one lateral convolution per backbone output is required")
        self.lateral_convs = nn.ModuleList(convs)

    def forward(self, img: torch.Tensor):
        bb = self.backbone
        dt = L.F32 if bb.compute_dtype == "fp32" else L.BF16
        outs = []
        for conv, (i, x_out, H, W) in zip(self.lateral_convs, bb.forward_tokens(img)):
            n = getattr(bb, f"norm{i}")
            outs.append(NormLateralFn.apply(x_out, n.weight, n.bias, conv.weight, conv.bias, H, W, dt, float(n.eps)))
        return outs
