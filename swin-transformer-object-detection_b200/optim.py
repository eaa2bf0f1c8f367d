"""Fused AdamW for the backbone's parameters (the step after backward).

The reference trains with ``AdamW(lr=1e-4, betas=(0.9, 0.999), weight_decay=0.05)`` and a paramwise rule that switches
weight decay off for every parameter whose name contains ``absolute_pos_embed``, ``relative_position_bias_table`` or
``norm`` (configs/swin/mask_rcnn_swin_tiny_patch4_window7_mstrain_480-800_adamw_1x_coco.py:64-67, applied by mmcv's
DefaultOptimizerConstructor); ``optimizer.step()`` is driven by ``DistOptimizerHook.after_train_iter``
(mmdet/utils/optimizer.py:22-33) and runs one small kernel chain per parameter tensor (189 tensors for Swin-T).

``FusedAdamW`` does the same arithmetic (torch.optim.AdamW, decoupled decay, bias-corrected moments) with one
``swin_adamw_step`` launch per 64 tensors, reading the gradients wherever they live (the all-reduce bucket views set
by ``BucketedGradAllReduce``) and refreshing, in the same pass, the bf16 shadow copies the tcgen05 GEMMs read — so the
operand cast of the next step is free and a captured CUDA graph keeps pointing at valid weights.  CUDA only.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch

from . import functional as F_
from . import ops

REFERENCE_CUSTOM_KEYS = {"absolute_pos_embed": 0.0, "relative_position_bias_table": 0.0, "norm": 0.0}


def paramwise_weight_decay(named_params: Iterable[Tuple[str, torch.nn.Parameter]], weight_decay: float,
                           custom_keys: Optional[Dict[str, float]] = None) -> List[Tuple[torch.nn.Parameter, float]]:
    """(parameter, weight decay) pairs under the reference's paramwise rule: the first custom key (longest first, then
    alphabetical — mmcv's order) that is a substring of the parameter name scales the base decay by its ``decay_mult``."""
    keys = REFERENCE_CUSTOM_KEYS if custom_keys is None else custom_keys
    order = sorted(sorted(keys.keys()), key=len, reverse=True)
    out = []
    for name, p in named_params:
        if not p.requires_grad:
            continue
        wd = weight_decay
        for k in order:
            if k in name:
                wd = weight_decay * float(keys[k])
                break
        out.append((p, wd))
    return out


class FusedAdamW:
    """``opt = FusedAdamW(model, lr=1e-4, weight_decay=0.05); loss.backward(); ddp.finish(); opt.step()``."""

    def __init__(self, module: torch.nn.Module, lr: float = 1e-4, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.05, custom_keys: Optional[Dict[str, float]] = None):
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        pairs = paramwise_weight_decay(module.named_parameters(), weight_decay, custom_keys)
        self.params = [p for p, _ in pairs]
        self.weight_decays = [wd for _, wd in pairs]
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32:
                raise RuntimeError("FusedAdamW needs fp32 CUDA parameters (there is no CPU fallback)")
        self.exp_avg = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.steps = 0

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0) -> None:
        self.steps += 1
        ps, gs, ms, vs, ws, wd = [], [], [], [], [], []
        for p, m, v, d in zip(self.params, self.exp_avg, self.exp_avg_sq, self.weight_decays):
            g = p.grad
            if g is None:
                continue
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
            hit = F_._W16.get(id(p))
            shadow = hit[2] if (hit is not None and hit[0]() is p and hit[2].device == p.device) else None
            ps.append(p.data); gs.append(g); ms.append(m); vs.append(v); ws.append(shadow); wd.append(d)
        ops.adamw_step(ps, gs, ms, vs, ws, wd, self.lr, self.betas[0], self.betas[1], self.eps, self.steps, grad_scale)

    def state_dict(self):
        return {"steps": self.steps, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr}

    def load_state_dict(self, sd) -> None:
        self.steps = int(sd["steps"]); self.lr = float(sd.get("lr", self.lr))
        for dst, src in zip(self.exp_avg, sd["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, sd["exp_avg_sq"]):
            dst.copy_(src)
