"""Fused AdamW for the backbone's parameters (the step after backward).

The reference trains with ``AdamW(lr=1e-4, betas=(0.9, 0.999), weight_decay=0.05)`` and a paramwise rule that switches
weight decay off for every parameter whose name contains ``absolute_pos_embed``, ``relative_position_bias_table`` or
``norm`` (configs/swin/mask_rcnn_swin_tiny_patch4_window7_mstrain_480-800_adamw_1x_coco.py:64-67, applied by mmcv's
DefaultOptimizerConstructor); ``optimizer.step()`` is driven by ``DistOptimizerHook.after_train_iter``
(mmdet/utils/optimizer.py:22-33) and runs one small kernel chain per parameter tensor (189 tensors for Swin-T).

``FusedAdamW`` is a ``torch.optim.Optimizer``: it has ``param_groups`` (per-group ``lr`` / ``betas`` / ``eps`` /
``weight_decay``, read afresh on every ``step()`` so mmcv's ``LrUpdaterHook`` -- warm-up + step schedule -- and any
torch ``lr_scheduler`` drive it), per-parameter ``state`` (``step``, ``exp_avg``, ``exp_avg_sq``) and the torch-format
``state_dict()`` / ``load_state_dict()``, so the ``optimizer`` entry of a reference checkpoint resumes here and
checkpoints saved here are readable by ``torch.optim.AdamW``.  The arithmetic is torch.optim.AdamW's (decoupled decay,
bias-corrected moments) done by one ``swin_adamw_step`` launch per 64 tensors, reading the gradients wherever they
live (the all-reduce bucket views set by ``BucketedGradAllReduce``) and refreshing, in the same pass, the bf16 shadow
copies the tcgen05 GEMMs read -- so the operand cast of the next step is free and a captured CUDA graph of the
forward/backward keeps pointing at valid weights.  CUDA only.

``step()`` must stay OUTSIDE a captured CUDA graph: the learning rate and the bias corrections are computed on the host
from ``group['lr']`` and the per-parameter step count and passed to the kernel by value.
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


def reference_param_groups(module: torch.nn.Module, weight_decay: float = 0.05,
                           custom_keys: Optional[Dict[str, float]] = None) -> List[dict]:
    """``param_groups`` under the reference's paramwise rule: one group per distinct weight decay (mmcv builds one group
    per parameter; the grouping does not change the arithmetic)."""
    by_wd: Dict[float, list] = {}
    for p, wd in paramwise_weight_decay(module.named_parameters(), weight_decay, custom_keys):
        by_wd.setdefault(wd, []).append(p)
    return [{"params": ps, "weight_decay": wd} for wd, ps in by_wd.items()]


class FusedAdamW(torch.optim.Optimizer):
    """``opt = FusedAdamW(model, lr=1e-4, weight_decay=0.05); loss.backward(); ddp.finish(); opt.step()``.

    ``params`` may be an ``nn.Module`` (the reference's paramwise weight-decay rule is applied to its named parameters),
    an iterable of parameters, or an iterable of ``param_groups`` dicts exactly as for ``torch.optim.AdamW``."""

    def __init__(self, params, lr: float = 1e-4, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.05, custom_keys: Optional[Dict[str, float]] = None):
        if isinstance(params, torch.nn.Module):
            params = reference_param_groups(params, weight_decay, custom_keys)
        defaults = dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps), weight_decay=float(weight_decay))
        super().__init__(params, defaults)
        for g in self.param_groups:
            for p in g["params"]:
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdamW needs fp32 CUDA parameters (there is no CPU fallback)")

    def _init_state(self, p) -> dict:
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)           # host scalar, as torch.optim.AdamW (non-capturable)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return st

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("FusedAdamW.step() must run outside CUDA-graph capture (lr and bias corrections are host values)")
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # one launch list per distinct (lr, betas, eps, step): normally one per param group
        batches: Dict[tuple, list] = {}
        for g in self.param_groups:
            lr, (b1, b2), eps, wd = float(g["lr"]), g["betas"], float(g["eps"]), float(g["weight_decay"])
            for p in g["params"]:
                grad = p.grad
                if grad is None:
                    continue
                if grad.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                if grad.dtype != torch.float32 or not grad.is_contiguous():
                    grad = grad.float().contiguous()
                st = self._init_state(p)
                st["step"] += 1
                hit = F_._W16.get(id(p))
                shadow = hit[2] if (hit is not None and hit[0]() is p and hit[2].device == p.device) else None
                key = (lr, float(b1), float(b2), eps, int(st["step"].item()))
                batches.setdefault(key, []).append((p, grad, st["exp_avg"], st["exp_avg_sq"], shadow, wd))
        for (lr, b1, b2, eps, step), items in batches.items():
            ps, gs, ms, vs, ws, wd = ([it[i] for it in items] for i in range(6))
            with torch.cuda.device(ps[0].device):
                ops.adamw_step([p.data for p in ps], gs, ms, vs, ws, wd, lr, b1, b2, eps, step, grad_scale)
        return loss
