"""Registration of the backbone under the reference's plug-in API.

The reference builds backbones with ``build_from_cfg(cfg, BACKBONES)`` where ``BACKBONES`` is the
mmcv Registry in ``mmdet/models/builder.py:6`` and the class is registered by the decorator at
``mmdet/models/backbones/swin_transformer.py:448``.  When mmdet is importable we register there
(``force=True`` replaces the stock class, so every ``configs/swin/*`` detector picks this one up
through ``custom_imports``); otherwise a local registry with the same ``register_module`` /
``build`` surface is used so tests and benchmarks can build from reference-style config dicts.
"""
from __future__ import annotations

from typing import Any, Dict


class LocalRegistry:
    """The subset of mmcv.utils.Registry the backbone path uses (mmdet/models/builder.py:15-39)."""

    def __init__(self, name: str):
        self.name = name
        self.module_dict: Dict[str, Any] = {}

    def register_module(self, name=None, force: bool = False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self.module_dict[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key: str):
        return self.module_dict.get(key)

    def build(self, cfg: Dict[str, Any]):
        cfg = dict(cfg)
        typ = cfg.pop("type")
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**cfg)


def _find_mmdet_registry():
    try:
        from mmdet.models.builder import BACKBONES as reg  # type: ignore
        return reg
    except Exception:
        return None


MMDET_BACKBONES = _find_mmdet_registry()
BACKBONES = LocalRegistry("backbone")


def register_backbone(cls):
    BACKBONES.register_module(force=True)(cls)
    if MMDET_BACKBONES is not None:
        MMDET_BACKBONES.register_module(force=True)(cls)
    return cls


def build_backbone(cfg: Dict[str, Any]):
    """Same contract as mmdet.models.builder.build_backbone (builder.py:37-39)."""
    return BACKBONES.build(cfg)
