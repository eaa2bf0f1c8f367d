"""swin_b200 — B200 (sm_100a) native implementation of the Swin backbone's shifted-window
attention path, drop-in for ``mmdet/models/backbones/swin_transformer.py`` of
AbdulHannanKhan/Swin-Transformer-Object-Detection.  Import as ``swin_b200``."""
from . import _lib  # noqa: F401
from .registry import BACKBONES, build_backbone  # noqa: F401
from .swin_transformer import (BasicLayer, DropPath, Mlp, PatchEmbed, PatchMerging, SwinTransformer,  # noqa: F401
                               SwinTransformerBlock, WindowAttention, window_partition, window_reverse)

__all__ = ["SwinTransformer", "SwinTransformerBlock", "WindowAttention", "PatchMerging", "BasicLayer", "PatchEmbed",
           "Mlp", "DropPath", "window_partition", "window_reverse", "BACKBONES", "build_backbone"]
