"""CPU oracle for the Swin shifted-window-attention path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch, *functional* restatement (plain torch-on-CPU / numpy index
arithmetic, no nn.Module, no roll/pad/permute chains) of what the reference computes in
``mmdet/models/backbones/swin_transformer.py``.  It exists only to check the CUDA path:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
  ``--impl reference`` legs may import it;
* the product package (``swin_b200``) never imports anything from ``oracle/`` and has no CPU
  fallback.

Pinning: the reference ships no golden vectors or known-answer tests for this path
(SURVEY.md §4, §8c), so the oracle is pinned against *outputs of the reference itself*:
``oracle/make_golden.py`` imports the unmodified reference file from ``/root/reference``
through five stub modules and writes seeded input/weight/output/gradient fixtures to
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against them
(bit-exact for the index ops, <=1e-5 rel-L2 for float), and when ``/root/reference`` is present
``tests/test_oracle_vs_reference.py`` compares live.

Every function cites the reference lines it restates (paths relative to /root/reference).
All tensors are torch CPU tensors so autograd provides the gradient oracle.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

REF = "mmdet/models/backbones/swin_transformer.py"


# --------------------------------------------------------------------------------------
# integer / index oracles (bit-exact class)
# --------------------------------------------------------------------------------------
def padded_hw(H: int, W: int, ws: int) -> Tuple[int, int]:
    """Hp, Wp of the zero-padded grid (REF:216-219, :371-372)."""
    return -(-H // ws) * ws, -(-W // ws) * ws


def window_partition_np(x: np.ndarray, ws: int) -> np.ndarray:
    """(B,Hp,Wp,C) -> (B*nW, ws, ws, C).  REF:41-53 stated as an explicit index map:
    window w = b*nW + wh*(Wp/ws) + ww ; element (i,j) <- pixel (wh*ws+i, ww*ws+j)."""
    B, Hp, Wp, C = x.shape
    nwh, nww = Hp // ws, Wp // ws
    out = np.empty((B * nwh * nww, ws, ws, C), dtype=x.dtype)
    for b in range(B):
        for wh in range(nwh):
            for ww in range(nww):
                w = (b * nwh + wh) * nww + ww
                out[w] = x[b, wh * ws:(wh + 1) * ws, ww * ws:(ww + 1) * ws, :]
    return out


def window_reverse_np(windows: np.ndarray, ws: int, Hp: int, Wp: int) -> np.ndarray:
    """(B*nW, ws, ws, C) -> (B,Hp,Wp,C).  REF:56-70 (inverse of window_partition)."""
    nwh, nww = Hp // ws, Wp // ws
    B = windows.shape[0] // (nwh * nww)
    C = windows.shape[-1]
    out = np.empty((B, Hp, Wp, C), dtype=windows.dtype)
    for b in range(B):
        for wh in range(nwh):
            for ww in range(nww):
                w = (b * nwh + wh) * nww + ww
                out[b, wh * ws:(wh + 1) * ws, ww * ws:(ww + 1) * ws, :] = windows[w]
    return out


def gather_index(H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """Closed form of pad -> roll(-s,-s) -> window_partition (REF:214-231).

    Returns int64 array (nW*ws*ws,) with, for every window-token slot of ONE image, the source
    token index ``hs*W + wsrc`` in the un-padded (H,W) grid, or -1 where the slot reads padding:
        hs = (wh*ws + i + s) mod Hp,  wsrc = (ww*ws + j + s) mod Wp,  valid iff hs<H and wsrc<W.
    """
    Hp, Wp = padded_hw(H, W, ws)
    nwh, nww = Hp // ws, Wp // ws
    wh, ww, i, j = np.meshgrid(np.arange(nwh), np.arange(nww), np.arange(ws), np.arange(ws), indexing="ij")
    hs = (wh * ws + i + shift) % Hp
    wsrc = (ww * ws + j + shift) % Wp
    idx = np.where((hs < H) & (wsrc < W), hs * W + wsrc, -1)
    return idx.reshape(-1).astype(np.int64)


_IDX_CACHE: Dict[tuple, torch.Tensor] = {}


def _cached(kind: str, key: tuple, device, make) -> torch.Tensor:
    """Index / mask tensors per (geometry, device), built once: the closed forms are pure functions of the geometry, and
    rebuilding a million-entry numpy index for every block would dominate the timing of the oracle when bench.py runs it
    on a GPU as the eager baseline."""
    k = (kind,) + key + (str(device),)
    t = _IDX_CACHE.get(k)
    if t is None:
        if len(_IDX_CACHE) > 256:
            _IDX_CACHE.clear()
        t = _IDX_CACHE[k] = make().to(device)
    return t


def shift_gather(x: torch.Tensor, H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """(B, H*W, C) -> (B*nW, ws*ws, C): REF:212-231 composite (pad/roll/partition) via gather_index."""
    B, L, C = x.shape
    idx = _cached("gather", (H, W, ws, shift), x.device, lambda: torch.from_numpy(gather_index(H, W, ws, shift)))
    xz = torch.cat([x, x.new_zeros(B, 1, C)], dim=1)          # slot L == the zero pad token
    sel = torch.where(idx < 0, torch.full_like(idx, L), idx)
    out = xz[:, sel, :]                                       # (B, nW*N, C)
    return out.reshape(-1, ws * ws, C)


def shift_scatter(xw: torch.Tensor, B: int, H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """(B*nW, ws*ws, C) -> (B, H*W, C): REF:236-249 composite (reverse / roll back / crop).
    Every valid pixel is written exactly once (the map is a permutation restricted to valid slots)."""
    C = xw.shape[-1]
    idx = _cached("gather", (H, W, ws, shift), xw.device, lambda: torch.from_numpy(gather_index(H, W, ws, shift)))
    flat = xw.reshape(B, -1, C)
    valid = idx >= 0
    out = xw.new_zeros(B, H * W, C)
    out[:, idx[valid], :] = flat[:, valid, :]
    return out


def region_id(p: np.ndarray, P: int, ws: int, shift: int) -> np.ndarray:
    """Label of REF:374-384's three slices along one axis of the padded grid:
    [0,P-ws) -> 0, [P-ws,P-s) -> 1, [P-s,P) -> 2."""
    return (p >= P - ws).astype(np.int64) + (p >= P - shift).astype(np.int64)


def shift_mask_np(H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """(nW, N, N) float32 additive mask of {0,-100}: REF:370-389 in closed form."""
    Hp, Wp = padded_hw(H, W, ws)
    nwh, nww = Hp // ws, Wp // ws
    wh, ww, i, j = np.meshgrid(np.arange(nwh), np.arange(nww), np.arange(ws), np.arange(ws), indexing="ij")
    reg = 3 * region_id(wh * ws + i, Hp, ws, shift) + region_id(ww * ws + j, Wp, ws, shift)
    reg = reg.reshape(nwh * nww, ws * ws)
    diff = reg[:, None, :] - reg[:, :, None]
    return np.where(diff != 0, np.float32(-100.0), np.float32(0.0)).astype(np.float32)


def relative_position_index_np(ws: int) -> np.ndarray:
    """(N,N) int64: (ri-rj+ws-1)*(2ws-1) + (ci-cj+ws-1).  REF:101-111."""
    r, c = np.divmod(np.arange(ws * ws), ws)
    return ((r[:, None] - r[None, :] + ws - 1) * (2 * ws - 1) + (c[:, None] - c[None, :] + ws - 1)).astype(np.int64)


def merge_index(H: int, W: int) -> np.ndarray:
    """PatchMerging gather (REF:281-293): (H2*W2, 4) source token index (or -1 for the odd pad),
    channel-block order [(r0,c0),(r1,c0),(r0,c1),(r1,c1)]."""
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    oh, ow = np.meshgrid(np.arange(H2), np.arange(W2), indexing="ij")
    out = np.empty((H2 * W2, 4), dtype=np.int64)
    for q, (dr, dc) in enumerate([(0, 0), (1, 0), (0, 1), (1, 1)]):
        r, c = 2 * oh + dr, 2 * ow + dc
        out[:, q] = np.where((r < H) & (c < W), r * W + c, -1).reshape(-1)
    return out


# --------------------------------------------------------------------------------------
# float oracles
# --------------------------------------------------------------------------------------
FAST_OPS = False     # bench.py's eager-GPU arm only: torch's fused layer_norm / gelu kernels (what the reference's nn.LayerNorm /
                     # nn.GELU launch) instead of the step-by-step restatements below


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.LayerNorm over the last dim, biased variance, eps 1e-5 (REF:185,191,269,549-553)."""
    if FAST_OPS:
        return F.layer_norm(x, (x.shape[-1],), w, b, eps)
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """nn.GELU() exact form (REF:23,34)."""
    if FAST_OPS:
        return F.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def window_attention(xw: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str, num_heads: int, ws: int,
                     mask: Optional[torch.Tensor] = None, qk_scale: Optional[float] = None) -> torch.Tensor:
    """WindowAttention.forward (REF:121-153) on (B_, N, C) windows.

    qkv columns are [q | k | v] x [head] x [head_dim] (REF:129); q is scaled before QK^T (REF:132);
    bias = table[index] (REF:135-138); window w uses mask[w mod nW] (REF:141-143)."""
    B_, N, C = xw.shape
    d = C // num_heads
    scale = qk_scale or d ** -0.5
    qkv = xw @ p[prefix + "qkv.weight"].t()
    if (prefix + "qkv.bias") in p:
        qkv = qkv + p[prefix + "qkv.bias"]
    q = qkv[..., 0 * C:1 * C].reshape(B_, N, num_heads, d).transpose(1, 2) * scale
    k = qkv[..., 1 * C:2 * C].reshape(B_, N, num_heads, d).transpose(1, 2)
    v = qkv[..., 2 * C:3 * C].reshape(B_, N, num_heads, d).transpose(1, 2)
    s = torch.einsum("bhid,bhjd->bhij", q, k)
    idx = _cached("relidx", (ws,), xw.device, lambda: torch.from_numpy(relative_position_index_np(ws)).reshape(-1))
    bias = p[prefix + "relative_position_bias_table"][idx].reshape(N, N, num_heads).permute(2, 0, 1)
    s = s + bias[None]
    if mask is not None:
        nW = mask.shape[0]
        widx = torch.arange(B_, device=xw.device) % nW
        s = s + mask[widx][:, None, :, :]
    a = torch.softmax(s, dim=-1)
    o = torch.einsum("bhij,bhjd->bihd", a, v).reshape(B_, N, C)
    return o @ p[prefix + "proj.weight"].t() + p[prefix + "proj.bias"]


def mlp(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """Mlp.forward (REF:32-38), dropout p=0."""
    h = gelu_erf(x @ p[prefix + "fc1.weight"].t() + p[prefix + "fc1.bias"])
    return h @ p[prefix + "fc2.weight"].t() + p[prefix + "fc2.bias"]


def swin_block(x: torch.Tensor, H: int, W: int, p: Dict[str, torch.Tensor], prefix: str, num_heads: int,
               ws: int, shift: int, drop_scale: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
               qk_scale: Optional[float] = None) -> torch.Tensor:
    """SwinTransformerBlock.forward (REF:198-255).  LN1 happens BEFORE the zero padding (REF:211
    then :218) so padded slots are exact zeros entering qkv.  ``drop_scale`` = per-sample
    (B,) multipliers for the two DropPath sites (REF:252-253); None = identity."""
    B, L, C = x.shape
    y = layer_norm(x, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"])
    yw = shift_gather(y, H, W, ws, shift)
    mask = _cached("mask", (H, W, ws, shift), x.device, lambda: torch.from_numpy(shift_mask_np(H, W, ws, shift))).to(x.dtype) if shift > 0 else None
    aw = window_attention(yw, p, prefix + "attn.", num_heads, ws, mask, qk_scale)
    a = shift_scatter(aw, B, H, W, ws, shift)
    if drop_scale is not None:
        a = a * drop_scale[0].reshape(B, 1, 1)
    x = x + a
    m = mlp(layer_norm(x, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"]), p, prefix + "mlp.")
    if drop_scale is not None:
        m = m * drop_scale[1].reshape(B, 1, 1)
    return x + m


def patch_merging(x: torch.Tensor, H: int, W: int, p: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """PatchMerging.forward (REF:271-298): zero-pad odd H/W, 2x2 gather in the order
    [(r0,c0),(r1,c0),(r0,c1),(r1,c1)], LayerNorm(4C), Linear(4C->2C, no bias)."""
    B, L, C = x.shape
    idx = _cached("merge", (H, W), x.device, lambda: torch.from_numpy(merge_index(H, W)))               # (L2, 4)
    xz = torch.cat([x, x.new_zeros(B, 1, C)], dim=1)
    sel = torch.where(idx < 0, torch.full_like(idx, L), idx)
    g = xz[:, sel.reshape(-1), :].reshape(B, idx.shape[0], 4 * C)
    g = layer_norm(g, p[prefix + "norm.weight"], p[prefix + "norm.bias"])
    return g @ p[prefix + "reduction.weight"].t()


def patch_embed(img: torch.Tensor, p: Dict[str, torch.Tensor], patch: int, patch_norm: bool) -> Tuple[torch.Tensor, int, int]:
    """PatchEmbed.forward (REF:429-445) as an explicit patch-unfold GEMM: zero-pad right/bottom to a
    multiple of the patch, (B,3,H,W) -> tokens (B, Wh*Ww, C) with optional LayerNorm."""
    B, Cin, H, W = img.shape
    img = F.pad(img, (0, (-W) % patch, 0, (-H) % patch))
    Wh, Ww = img.shape[2] // patch, img.shape[3] // patch
    cols = img.reshape(B, Cin, Wh, patch, Ww, patch).permute(0, 2, 4, 1, 3, 5).reshape(B, Wh * Ww, Cin * patch * patch)
    wmat = p["patch_embed.proj.weight"].reshape(p["patch_embed.proj.weight"].shape[0], -1)
    tok = cols @ wmat.t() + p["patch_embed.proj.bias"]
    if patch_norm:
        tok = layer_norm(tok, p["patch_embed.norm.weight"], p["patch_embed.norm.bias"])
    return tok, Wh, Ww


def drop_path_rates(drop_path_rate: float, depths: Sequence[int]) -> List[float]:
    """Stochastic-depth schedule linspace(0, rate, sum(depths)) (REF:525)."""
    return [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]


def backbone_forward(img: torch.Tensor, p: Dict[str, torch.Tensor], embed_dim: int = 96,
                     depths: Sequence[int] = (2, 2, 6, 2), num_heads: Sequence[int] = (3, 6, 12, 24),
                     window_size: int = 7, patch_size: int = 4, patch_norm: bool = True,
                     out_indices: Sequence[int] = (0, 1, 2, 3), qk_scale: Optional[float] = None,
                     drop_scales: Optional[List[Tuple[torch.Tensor, torch.Tensor]]] = None) -> Tuple[torch.Tensor, ...]:
    """SwinTransformer.forward (REF:600-625), dropout 0.  ``drop_scales`` optionally gives the host-drawn DropPath
    multipliers per block in execution order.  ape=True (REF:604-607) is selected by the presence of
    ``absolute_pos_embed`` (1, C, Hpre, Wpre) in ``p``: bicubic-interpolated to the token grid and added before the first block."""
    x, H, W = patch_embed(img, p, patch_size, patch_norm)
    if "absolute_pos_embed" in p:
        pos = torch.nn.functional.interpolate(p["absolute_pos_embed"], size=(H, W), mode="bicubic")      # REF:606
        x = x + pos.flatten(2).transpose(1, 2)
    outs = []
    blk_no = 0
    for s, depth in enumerate(depths):
        C = embed_dim * 2 ** s
        for b in range(depth):
            shift = 0 if b % 2 == 0 else window_size // 2        # REF:346
            ds = drop_scales[blk_no] if drop_scales is not None else None
            x = swin_block(x, H, W, p, f"layers.{s}.blocks.{b}.", num_heads[s], window_size, shift, ds, qk_scale)
            blk_no += 1
        if s in out_indices:
            o = layer_norm(x, p[f"norm{s}.weight"], p[f"norm{s}.bias"])
            outs.append(o.reshape(-1, H, W, C).permute(0, 3, 1, 2).contiguous())   # REF:618-623
        if s < len(depths) - 1:
            x = patch_merging(x, H, W, p, f"layers.{s}.downsample.")
            H, W = (H + 1) // 2, (W + 1) // 2
    return tuple(outs)


# --------------------------------------------------------------------------------------
# seeded weights / inputs shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------------------
def param_shapes(embed_dim: int, depths: Sequence[int], num_heads: Sequence[int], window_size: int = 7,
                 patch_size: int = 4, in_chans: int = 3, mlp_ratio: float = 4.0,
                 out_indices: Sequence[int] = (0, 1, 2, 3)) -> Dict[str, Tuple[int, ...]]:
    """state_dict parameter names/shapes of the reference backbone, in module order."""
    sh: Dict[str, Tuple[int, ...]] = {}
    sh["patch_embed.proj.weight"] = (embed_dim, in_chans, patch_size, patch_size)
    sh["patch_embed.proj.bias"] = (embed_dim,)
    sh["patch_embed.norm.weight"] = (embed_dim,)
    sh["patch_embed.norm.bias"] = (embed_dim,)
    for s, depth in enumerate(depths):
        C = embed_dim * 2 ** s
        hid = int(C * mlp_ratio)
        for b in range(depth):
            pre = f"layers.{s}.blocks.{b}."
            sh[pre + "norm1.weight"] = (C,)
            sh[pre + "norm1.bias"] = (C,)
            sh[pre + "attn.relative_position_bias_table"] = ((2 * window_size - 1) ** 2, num_heads[s])
            sh[pre + "attn.qkv.weight"] = (3 * C, C)
            sh[pre + "attn.qkv.bias"] = (3 * C,)
            sh[pre + "attn.proj.weight"] = (C, C)
            sh[pre + "attn.proj.bias"] = (C,)
            sh[pre + "norm2.weight"] = (C,)
            sh[pre + "norm2.bias"] = (C,)
            sh[pre + "mlp.fc1.weight"] = (hid, C)
            sh[pre + "mlp.fc1.bias"] = (hid,)
            sh[pre + "mlp.fc2.weight"] = (C, hid)
            sh[pre + "mlp.fc2.bias"] = (C,)
        if s < len(depths) - 1:
            sh[f"layers.{s}.downsample.reduction.weight"] = (2 * C, 4 * C)
            sh[f"layers.{s}.downsample.norm.weight"] = (4 * C,)
            sh[f"layers.{s}.downsample.norm.bias"] = (4 * C,)
    for i in out_indices:
        sh[f"norm{i}.weight"] = (embed_dim * 2 ** i,)
        sh[f"norm{i}.bias"] = (embed_dim * 2 ** i,)
    return sh


def seeded_params(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, std: float = 0.02,
                  noisy: bool = True, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic weights from a numpy Generator (independent of torch's RNG stream/version).
    ``noisy``: biases / LN affine / bias tables get sigma=0.3-ish noise so padded-key, bias and
    affine effects are visible in parity checks; else reference-style init."""
    rng = np.random.default_rng(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in shapes.items():
        if name.endswith("norm.weight") or ".norm1.weight" in name or ".norm2.weight" in name or \
                (name.startswith("norm") and name.endswith(".weight")):
            a = 1.0 + (0.2 * rng.standard_normal(shape) if noisy else 0.0)
        elif name.endswith(".bias"):
            a = 0.3 * rng.standard_normal(shape) if noisy else np.zeros(shape)
        elif name.endswith("relative_position_bias_table"):
            a = (0.3 if noisy else std) * rng.standard_normal(shape)
        elif name == "patch_embed.proj.weight":
            a = rng.standard_normal(shape) / math.sqrt(shape[1] * shape[2] * shape[3])
        else:
            fan_in = shape[-1]
            a = rng.standard_normal(shape) * (1.0 / math.sqrt(fan_in) if noisy else std)
        out[name] = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), shape))).to(dtype)
    return out


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in float64 (the parity metric of SURVEY.md §8c)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)
