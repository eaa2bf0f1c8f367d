"""Host-side mirror of the reference's Swin backbone interface, running on the sm_100a kernels.

Same class names, constructor arguments, ``forward`` signatures, sub-module names and
``state_dict`` keys as ``mmdet/models/backbones/swin_transformer.py`` (cited as REF:line), so
reference checkpoints load unchanged and every ``configs/swin/*`` detector can use it as a
drop-in.  The arithmetic is NOT the reference's op chain: each block is one fused
autograd.Function over the C-ABI kernels (functional.py).  There is no CPU path — calling
``forward`` on CPU tensors raises.

Extra, optional constructor argument: ``compute_dtype`` ("bf16" default | "fp32").
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint as cp

from . import _lib as L
from . import ops
from .functional import BlockLink, MlpFn, OutNormFn, OutNormMergeFn, PatchEmbedFn, PatchMergingFn, SwinBlockFn, WindowAttentionFn
from .registry import register_backbone

_STAGE_LINK = os.environ.get("SWIN_STAGE_LINK", "1") != "0"      # A/B knob: 0 = the PatchMerging dY cast runs as its own scale_cast pass


def _dt_code(compute_dtype: Optional[str]) -> int:
    v = (compute_dtype or os.environ.get("SWIN_B200_DTYPE", "bf16")).lower()
    if v in ("bf16", "bfloat16"):
        return L.BF16
    if v in ("fp32", "float32", "f32"):
        return L.F32
    raise ValueError(f"compute_dtype must be 'bf16' or 'fp32', got {compute_dtype!r}")


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _need_cuda(t: torch.Tensor, who: str):
    if not t.is_cuda:
        raise RuntimeError(f"{who}: swin_b200 runs only on CUDA (sm_100a) tensors; there is no CPU fallback")


def window_partition(x: torch.Tensor, window_size: int) -> torch.Tensor:
    """(B, H, W, C) -> (num_windows*B, window_size, window_size, C).  REF:41-53 (bit-exact copy kernel)."""
    _need_cuda(x, "window_partition")
    return ops.window_partition(x.contiguous(), window_size)


def window_reverse(windows: torch.Tensor, window_size: int, H: int, W: int) -> torch.Tensor:
    """(num_windows*B, window_size, window_size, C) -> (B, H, W, C).  REF:56-70."""
    _need_cuda(windows, "window_reverse")
    return ops.window_reverse(windows.contiguous(), window_size, H, W)


def _canonical_grid(mask: torch.Tensor, ws: int, shift: int):
    """(nwh, nww) if ``mask`` is bit-for-bit the canonical SW-MSA mask (REF:370-389) of some window grid with nwh*nww ==
    mask.shape[0], else (0, 0).  Masks made by BasicLayer.attn_mask carry the answer as the ``_swin_canon`` attribute.  A
    mask that arrives from elsewhere (e.g. built by reference code) is compared against the candidate grids on the device;
    the verdict is remembered ON THE TENSOR OBJECT together with its version counter -- never under its address, which the
    caching allocator hands to the next mask of a different grid (3x4 vs 4x3 windows in multi-scale training) -- so a
    rebuilt mask is verified again and an in-place edit invalidates the verdict."""
    tag = getattr(mask, "_swin_canon", None)
    if tag is not None:
        return tag
    if ws != 7 or shift != 3 or mask.dim() != 3 or mask.shape[1] != ws * ws:
        return (0, 0)
    seen = getattr(mask, "_swin_canon_checked", None)
    if seen is not None and seen[0] == mask._version:
        return seen[1]
    if torch.cuda.is_current_stream_capturing():
        return (0, 0)                      # cannot compare (needs a host read) while capturing: honour the tensor
    nW, verdict = mask.shape[0], (0, 0)
    m32 = mask.detach()
    if m32.dtype != torch.float32 or not m32.is_contiguous():
        m32 = m32.float().contiguous()
    for nwh in range(1, nW + 1):
        if nW % nwh:
            continue
        nww = nW // nwh
        cand = ops.shift_mask(nwh * ws, nww * ws, ws, shift, mask.device)
        if torch.equal(cand, m32):
            verdict = (nwh, nww)
            break
    try:
        mask._swin_canon_checked = (mask._version, verdict)
    except Exception:
        pass
    return verdict


class DropPath(nn.Module):
    """Stochastic depth per sample (timm semantics used at REF:190,252,253).  ``sample_scale`` returns the
    (B,) multiplier floor(keep + U[0,1)) / keep that the fused residual epilogues apply; draws one uniform per
    sample per call, in the reference's call order."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self._pending = []          # scales pre-drawn for this forward by SwinTransformer (one batched draw per step)

    def sample_scale(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        if self.drop_prob == 0.0 or not self.training:
            return None
        if self._pending:
            s = self._pending.pop(0)
            if s.shape[0] == x.shape[0] and s.device == x.device:
                return s
        keep = 1.0 - self.drop_prob
        r = keep + torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=torch.float32, device=x.device)
        return (r.floor_() / keep).reshape(-1).contiguous()

    def forward(self, x):
        s = self.sample_scale(x)
        return x if s is None else x * s.view((-1,) + (1,) * (x.ndim - 1)).to(x.dtype)

    def extra_repr(self):
        return f"p={self.drop_prob}"


class Mlp(nn.Module):
    """fc1 -> GELU(erf) -> fc2 (REF:20-38).  Inside a block the fused path uses these parameters directly;
    called standalone it runs the same GEMM kernels."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0,
                 compute_dtype: Optional[str] = None):
        super().__init__()
        if act_layer is not nn.GELU:
            raise NotImplementedError("swin_b200 fuses exact-erf GELU; other activations are not supported")
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)
        self.drop_rate = drop

        self._dt = _dt_code(compute_dtype)

    def forward(self, x):
        """x (..., in_features) -> (..., out_features).  REF:32-38 (dropout p = 0)."""
        _need_cuda(x, "Mlp")
        if self.training and self.drop_rate > 0:
            raise NotImplementedError("dropout > 0 is not implemented in the fused kernels")
        return MlpFn.apply(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self._dt)


class WindowAttention(nn.Module):
    """W-MSA / SW-MSA with relative position bias (REF:73-153); same parameters/buffers."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
                 compute_dtype: Optional[str] = None):
        super().__init__()
        self.dim = dim
        self.window_size = _pair(window_size)
        if self.window_size[0] != self.window_size[1]:
            raise NotImplementedError("square windows only")
        self.num_heads = num_heads
        if dim % num_heads or dim // num_heads != 32:
            raise NotImplementedError("swin_b200 kernels are specialised for head_dim 32 (every Swin variant)")
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        ws = self.window_size[0]
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) ** 2, num_heads))
        r = torch.arange(ws * ws) // ws
        c = torch.arange(ws * ws) % ws
        idx = (r[:, None] - r[None, :] + ws - 1) * (2 * ws - 1) + (c[:, None] - c[None, :] + ws - 1)
        self.register_buffer("relative_position_index", idx.long())
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)
        self._drops = (attn_drop, proj_drop)
        self._dt = _dt_code(compute_dtype)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)

    def forward(self, x, mask=None):
        """x: (num_windows*B, N, C); mask: (num_windows, N, N) additive or None."""
        _need_cuda(x, "WindowAttention")
        if self.training and any(p > 0 for p in self._drops):
            raise NotImplementedError("attention/projection dropout > 0 is not implemented in the fused kernels")
        mask_nz = None
        canon = (0, 0)
        if mask is not None:
            canon = _canonical_grid(mask, self.window_size[0], self.window_size[0] // 2)    # on the caller's tensor object
            mask_nz = getattr(mask, "_swin_nz", None)
            mask = mask.detach().float().contiguous()
            if mask_nz is None:
                mask_nz = ops.mask_nonzero(mask)
        return WindowAttentionFn.apply(x, self.relative_position_bias_table, self.qkv.weight, self.qkv.bias,
                                       self.proj.weight, self.proj.bias, mask, mask_nz, canon, self.window_size[0], self.num_heads,
                                       float(self.scale), self._dt, not torch.is_grad_enabled())


class SwinTransformerBlock(nn.Module):
    """REF:156-255.  forward(x, mask_matrix) with self.H / self.W set by the caller (REF:207, :392)."""

    def __init__(self, dim, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 compute_dtype: Optional[str] = None):
        super().__init__()
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("swin_b200 fuses nn.LayerNorm; other norm layers are not supported")
        assert 0 <= shift_size < window_size, "shift_size must in 0-window_size"
        self.dim, self.num_heads, self.window_size, self.shift_size, self.mlp_ratio = dim, num_heads, window_size, shift_size, mlp_ratio
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, _pair(window_size), num_heads, qkv_bias, qk_scale, attn_drop, drop, compute_dtype)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), act_layer=act_layer, drop=drop, compute_dtype=compute_dtype)
        self.H = None
        self.W = None
        self._drop = drop
        self._dt = _dt_code(compute_dtype)

    def forward(self, x, mask_matrix, recv=None, send=None):
        """Reference signature ``forward(x, mask_matrix)`` (REF:198).  ``recv`` / ``send`` are the backward hand-over links to
        the next / previous block of the stage (functional.BlockLink), passed by BasicLayer.forward only."""
        _need_cuda(x, "SwinTransformerBlock")
        B, Lx, C = x.shape
        H, W = self.H, self.W
        assert Lx == H * W, "input feature has wrong size"
        if self.training and (self._drop > 0 or any(p > 0 for p in self.attn._drops)):
            raise NotImplementedError("dropout > 0 is not implemented in the fused kernels")
        mask = mask_nz = None
        canon = (0, 0)
        if self.shift_size > 0:
            if mask_matrix is None:
                raise ValueError("shifted block needs mask_matrix")
            mask = mask_matrix.detach().float().contiguous()
            mask_nz = getattr(mask_matrix, "_swin_nz", None)     # set by BasicLayer.attn_mask (cached per geometry)
            if mask_nz is None:
                mask_nz = ops.mask_nonzero(mask)
            # the mask built by BasicLayer.attn_mask is the canonical one: the kernel evaluates it in closed form
            canon = _canonical_grid(mask_matrix, self.window_size, self.shift_size)
        s1 = s2 = None
        if isinstance(self.drop_path, DropPath):
            s1 = self.drop_path.sample_scale(x)      # attention-branch draw first, then MLP (REF:252-253)
            s2 = self.drop_path.sample_scale(x)
        a, m = self.attn, self.mlp
        return SwinBlockFn.apply(x, self.norm1.weight, self.norm1.bias, a.relative_position_bias_table, a.qkv.weight,
                                 a.qkv.bias, a.proj.weight, a.proj.bias, self.norm2.weight, self.norm2.bias,
                                 m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, mask, mask_nz, s1, s2,
                                 H, W, self.window_size, self.shift_size, self.num_heads, float(a.scale), self._dt,
                                 float(self.norm1.eps), canon, recv, send, not torch.is_grad_enabled())


class PatchMerging(nn.Module):
    """REF:258-298: 2x2 gather (+ zero pad of odd H/W) -> LayerNorm(4C) -> Linear(4C, 2C, bias=False)."""

    def __init__(self, dim, norm_layer=nn.LayerNorm, compute_dtype: Optional[str] = None):
        super().__init__()
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("swin_b200 fuses nn.LayerNorm; other norm layers are not supported")
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)
        self._dt = _dt_code(compute_dtype)

    def forward(self, x, H, W):
        _need_cuda(x, "PatchMerging")
        B, Lx, C = x.shape
        assert Lx == H * W, "input feature has wrong size"
        return PatchMergingFn.apply(x, self.norm.weight, self.norm.bias, self.reduction.weight, H, W, self._dt,
                                    float(self.norm.eps))


class BasicLayer(nn.Module):
    """One stage (REF:301-402): blocks alternate shift 0 / window_size//2, optional downsample."""

    def __init__(self, dim, depth, num_heads, window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False,
                 compute_dtype: Optional[str] = None):
        super().__init__()
        self.window_size = window_size
        self.shift_size = window_size // 2
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList(
            SwinTransformerBlock(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2, mlp_ratio, qkv_bias,
                                 qk_scale, drop, attn_drop, drop_path[i] if isinstance(drop_path, (list, tuple)) else drop_path,
                                 norm_layer=norm_layer, compute_dtype=compute_dtype)
            for i in range(depth))
        self.downsample = downsample(dim=dim, norm_layer=norm_layer, compute_dtype=compute_dtype) if downsample is not None else None
        self._mask_cache = {}

    def attn_mask(self, H: int, W: int, device) -> torch.Tensor:
        """SW-MSA mask (nW, N, N) fp32 of {0, -100} (REF:370-389), built by the bit-exact mask kernel."""
        key = (H, W, str(device))
        m = self._mask_cache.get(key)
        if m is None:
            m = ops.shift_mask(H, W, self.window_size, self.shift_size, device)
            m._swin_nz = ops.mask_nonzero(m)      # per-window "mask is not all-zero" flags for the attention kernels
            m._swin_canon = (-(-H // self.window_size), -(-W // self.window_size))   # canonical: closed form allowed
            if len(self._mask_cache) > 16:
                self._mask_cache.clear()
            self._mask_cache[key] = m
        return m

    def forward(self, x, H, W, out_norm=None, stage_send=None):
        """Reference signature ``forward(x, H, W) -> (x, H, W, x_down, Wh, Ww)`` (REF:362, :397-402).  With ``out_norm`` (the
        backbone's norm{i} module; SwinTransformer.forward passes it) the first element is already norm{i}(x) in NCHW:
        the output norm and the downsample then form one autograd node (functional.OutNormMergeFn).  ``stage_send``: the
        previous stage's link (its OutNormMergeFn takes the compute-dtype copy of its incoming gradient from this stage's first
        LN1 backward); this stage's own link is left in ``self.stage_link`` for the next one."""
        _need_cuda(x, "BasicLayer")
        mask = self.attn_mask(H, W, x.device)
        # backward hand-over between consecutive blocks (functional.BlockLink); not under activation checkpointing, whose
        # recomputation would run the blocks' forwards out of order
        link = torch.is_grad_enabled() and not self.use_checkpoint
        send = stage_send if link else None
        self.stage_link = None
        for bi, blk in enumerate(self.blocks):
            blk.H, blk.W = H, W
            if self.use_checkpoint:
                x = cp.checkpoint(blk, x, mask, use_reentrant=False)
            else:
                recv = BlockLink() if (link and bi + 1 < len(self.blocks)) else None
                x = blk(x, mask, recv, send)
                send = recv
        if self.downsample is not None:
            if out_norm is not None:
                ds = self.downsample
                if link and ds._dt == L.BF16 and _STAGE_LINK:
                    self.stage_link = BlockLink()
                out, x_down = OutNormMergeFn.apply(x, out_norm.weight, out_norm.bias, ds.norm.weight, ds.norm.bias,
                                                   ds.reduction.weight, H, W, ds._dt, float(out_norm.eps), float(ds.norm.eps),
                                                   self.stage_link)
                return out, H, W, x_down, (H + 1) // 2, (W + 1) // 2
            x_down = self.downsample(x, H, W)
            return x, H, W, x_down, (H + 1) // 2, (W + 1) // 2
        if out_norm is not None:
            return OutNormFn.apply(x, out_norm.weight, out_norm.bias, H, W, float(out_norm.eps)), H, W, x, H, W
        return x, H, W, x, H, W


class PatchEmbed(nn.Module):
    """REF:405-445: zero-pad to the patch multiple, patch x patch / patch conv, optional LayerNorm.  Runs as a
    patch-unfold gather + the GEMM kernels + the LayerNorm kernel; ``tokens()`` emits the (B, Wh*Ww, C) layout the
    block stack consumes, ``forward()`` keeps the reference's (B, C, Wh, Ww) return."""

    def __init__(self, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None, compute_dtype: Optional[str] = None):
        super().__init__()
        self.patch_size = _pair(patch_size)
        if self.patch_size[0] != self.patch_size[1]:
            raise NotImplementedError("square patches only")
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None
        self._dt = _dt_code(compute_dtype)

    def tokens(self, x):
        _need_cuda(x, "PatchEmbed")
        B, _, H, W = x.shape
        p = self.patch_size[0]
        nw, nb = (self.norm.weight, self.norm.bias) if self.norm is not None else (None, None)
        eps = float(self.norm.eps) if self.norm is not None else 1e-5
        t = PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, nw, nb, p, self._dt, eps)
        return t, -(-H // p), -(-W // p)

    def forward(self, x):
        t, Wh, Ww = self.tokens(x)
        return t.transpose(1, 2).reshape(-1, self.embed_dim, Wh, Ww)


@register_backbone
class SwinTransformer(nn.Module):
    """Swin backbone, REF:448-630.  Constructor arguments and defaults as REF:478-497 (+ compute_dtype)."""

    def __init__(self, pretrain_img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 out_indices=(0, 1, 2, 3), frozen_stages=-1, use_checkpoint=False, compute_dtype: Optional[str] = None):
        super().__init__()
        self.pretrain_img_size = pretrain_img_size
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.ape = ape
        self.patch_norm = patch_norm
        self.out_indices = out_indices
        self.frozen_stages = frozen_stages
        self.compute_dtype = "fp32" if _dt_code(compute_dtype) == L.F32 else "bf16"
        self.patch_embed = PatchEmbed(patch_size, in_chans, embed_dim, norm_layer if patch_norm else None, compute_dtype)
        if ape:
            pis, ps = _pair(pretrain_img_size), _pair(patch_size)
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, embed_dim, pis[0] // ps[0], pis[1] // ps[1]))
            nn.init.trunc_normal_(self.absolute_pos_embed, std=0.02)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]       # REF:525
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i), depth=depths[i], num_heads=num_heads[i], window_size=window_size,
                mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer=norm_layer,
                downsample=PatchMerging if i < self.num_layers - 1 else None, use_checkpoint=use_checkpoint,
                compute_dtype=compute_dtype))
        self.num_features = [int(embed_dim * 2 ** i) for i in range(self.num_layers)]
        for i in out_indices:
            self.add_module(f"norm{i}", norm_layer(self.num_features[i]))
        self._freeze_stages()

    def _freeze_stages(self):
        """REF:557-572."""
        if self.frozen_stages >= 0:
            self.patch_embed.eval()
            for p in self.patch_embed.parameters():
                p.requires_grad = False
        if self.frozen_stages >= 1 and self.ape:
            self.absolute_pos_embed.requires_grad = False
        if self.frozen_stages >= 2:
            self.pos_drop.eval()
            for i in range(0, self.frozen_stages - 1):
                self.layers[i].eval()
                for p in self.layers[i].parameters():
                    p.requires_grad = False

    def init_weights(self, pretrained=None):
        """REF:574-598: Linear trunc-normal(0.02) + zero bias, LayerNorm 1/0; a str loads a checkpoint."""
        def _init(m):
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        if pretrained is None or isinstance(pretrained, str):
            self.apply(_init)
            if isinstance(pretrained, str):
                from .checkpoint import load_checkpoint
                load_checkpoint(self, pretrained, strict=False)
        else:
            raise TypeError("pretrained must be a str or None")

    def _predraw_drop_path(self, B: int, device) -> None:
        """All stochastic-depth multipliers of one forward from ONE uniform draw: (2 per block with rate > 0, B) values
        floor(keep + U) / keep, handed to the blocks' DropPath modules in execution order (attention branch, then MLP).
        Same distribution as the reference's per-call draws (REF:190,252-253) with 4 launches per step instead of ~90."""
        mods = [blk.drop_path for layer in self.layers for blk in layer.blocks
                if isinstance(blk.drop_path, DropPath) and blk.drop_path.drop_prob > 0.0 and blk.drop_path.training]
        for m in mods:
            m._pending = []
        if not mods:
            return
        rates = tuple(m.drop_prob for m in mods)
        cache = getattr(self, "_keep_cache", None)
        if cache is None or cache[0] != (rates, str(device)):       # built once, outside any CUDA-graph capture
            keep = torch.tensor([1.0 - p for p in rates for _ in range(2)], dtype=torch.float32, device=device)
            self._keep_cache = ((rates, str(device)), keep)
        keep = self._keep_cache[1]
        r = torch.rand((2 * len(mods), B), dtype=torch.float32, device=device)
        scales = ((r + keep[:, None]).floor_() / keep[:, None]).contiguous()
        for i, m in enumerate(mods):
            m._pending = [scales[2 * i], scales[2 * i + 1]]

    def forward_tokens(self, x, apply_out_norm=False):
        """The block stack without the output norms: [(stage index, tokens (B, H*W, C) fp32, H, W)] for ``out_indices``
        (REF:600-617); ``swin_b200.fpn`` feeds FPN laterals from it directly.  ``forward`` calls it with
        ``apply_out_norm=True``: each stage then applies norm{i} + NCHW itself (fused with its downsample's backward) and
        the second tuple element is the finished (B, C, H, W) output."""
        _need_cuda(x, "SwinTransformer")
        if self.training and not any(layer.use_checkpoint for layer in self.layers):
            # (activation checkpointing re-runs the blocks under the saved RNG state, which only per-call draws reproduce)
            self._predraw_drop_path(x.shape[0], x.device)
        x, Wh, Ww = self.patch_embed.tokens(x)
        if self.ape:
            pos = F.interpolate(self.absolute_pos_embed, size=(Wh, Ww), mode="bicubic")
            x = x + pos.flatten(2).transpose(1, 2)
        x = self.pos_drop(x)
        outs = []
        for i, layer in enumerate(self.layers):
            n = getattr(self, f"norm{i}") if (apply_out_norm and i in self.out_indices) else None
            prev_link = getattr(self.layers[i - 1], "stage_link", None) if i > 0 else None
            x_out, H, W, x, Wh, Ww = layer(x, Wh, Ww, stage_send=prev_link) if n is None else layer(x, Wh, Ww, out_norm=n, stage_send=prev_link)
            if i in self.out_indices:
                outs.append((i, x_out, H, W))
        return outs

    def forward(self, x):
        return tuple(o for _, o, _, _ in self.forward_tokens(x, apply_out_norm=True))

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        return self
