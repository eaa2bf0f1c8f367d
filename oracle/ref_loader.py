"""Load the UNMODIFIED reference backbone file from /root/reference.  TEST INFRASTRUCTURE ONLY.

The reference (``mmdet/models/backbones/swin_transformer.py``) needs ``timm``, ``mmcv_custom`` and
``mmdet`` at import time (lines 13-17); none is installed here, so five stub modules are seeded
into ``sys.modules`` and the file is imported by path.  Nothing from the
reference is copied: the file is executed from where it lies.  Not available on the GPU box
(``/root/reference`` does not travel) -> ``available()`` is False there and callers skip.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("SWIN_REFERENCE_ROOT", "/root/reference")
REF_FILE = os.path.join(REF_ROOT, "mmdet", "models", "backbones", "swin_transformer.py")
_MOD = None


def available() -> bool:
    return os.path.isfile(REF_FILE)


class _DropPath(torch.nn.Module):
    """timm.models.layers.DropPath semantics (un-vendored dependency of the reference, line 13):
    identity in eval or p==0, else x/keep * floor(keep + U[0,1)) with one uniform per sample."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        r = keep + torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device)
        return x.div(keep) * r.floor_()


class _Registry:
    def __init__(self):
        self.module_dict = {}

    def register_module(self, *a, **k):
        def deco(cls):
            self.module_dict[cls.__name__] = cls
            return cls
        return deco


def load():
    """Return the reference module object (cached)."""
    global _MOD
    if _MOD is not None:
        return _MOD
    if not available():
        raise FileNotFoundError(REF_FILE)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    saved = {k: sys.modules.get(k) for k in
             ("timm", "timm.models", "timm.models.layers", "mmcv_custom", "mmdet", "mmdet.utils",
              "mmdet.models", "mmdet.models.builder", "mmdet.models.backbones")}
    try:
        mod("timm"); mod("timm.models")
        mod("timm.models.layers", DropPath=_DropPath, to_2tuple=lambda v: v if isinstance(v, tuple) else (v, v),
            trunc_normal_=torch.nn.init.trunc_normal_)
        mod("mmcv_custom", load_checkpoint=lambda *a, **k: None)
        mod("mmdet"); mod("mmdet.utils", get_root_logger=lambda *a, **k: None)
        mod("mmdet.models"); mod("mmdet.models.builder", BACKBONES=_Registry())
        mod("mmdet.models.backbones")
        name = "mmdet.models.backbones.swin_transformer"
        spec = importlib.util.spec_from_file_location(name, REF_FILE)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        _MOD = m
    finally:
        # do not leave fake mmdet/timm packages visible to the product's registry probing
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.modules.pop("mmdet.models.backbones.swin_transformer", None)
    return _MOD
