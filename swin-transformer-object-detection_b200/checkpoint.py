"""Checkpoint loading for the backbone: same fix-ups as the reference's Swin-aware loader
(``mmcv_custom/checkpoint.py:286-356``), written against plain ``torch.load`` (no mmcv):

* accepts ``{'state_dict': ...}``, ``{'model': ...}`` (ImageNet Swin releases) or a bare state dict (:316-322);
* strips a ``module.`` prefix (:324-325) and keeps only the ``encoder.`` branch of MoBY checkpoints (:328-329);
* reshapes ``absolute_pos_embed`` from (1, L, C) to (1, C, H, W) when the sizes agree (:332-339);
* bicubically resizes every ``relative_position_bias_table`` whose window size differs (:342-352);
* classifier-head keys that the backbone does not have are ignored (``strict=False``), as are the
  ``relative_position_index`` / ``attn_mask`` buffers of ImageNet checkpoints.
"""
from __future__ import annotations

import logging
from typing import Dict

import torch
import torch.nn.functional as F

_log = logging.getLogger("swin_b200")


def adapt_state_dict(model: torch.nn.Module, state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Apply the reference loader's key / shape fix-ups and return a dict ready for ``load_state_dict``."""
    sd = dict(state_dict)
    if sd and next(iter(sd)).startswith("module."):
        sd = {k[7:]: v for k, v in sd.items()}
    if sd and sorted(sd)[0].startswith("encoder"):
        sd = {k.replace("encoder.", ""): v for k, v in sd.items() if k.startswith("encoder.")}
    own = model.state_dict()
    ape = sd.get("absolute_pos_embed")
    if ape is not None and "absolute_pos_embed" in own and ape.dim() == 3:
        n1, l, c1 = ape.shape
        n2, c2, h, w = own["absolute_pos_embed"].shape
        if n1 != n2 or c1 != c2 or l != h * w:
            _log.warning("Error in loading absolute_pos_embed, pass")
            sd.pop("absolute_pos_embed")
        else:
            sd["absolute_pos_embed"] = ape.view(n2, h, w, c2).permute(0, 3, 1, 2).contiguous()
    for key in [k for k in sd if "relative_position_bias_table" in k]:
        if key not in own:
            continue
        pre, cur = sd[key], own[key]
        (l1, h1), (l2, h2) = pre.shape, cur.shape
        if h1 != h2:
            _log.warning("Error in loading %s, pass", key)
            sd.pop(key)
        elif l1 != l2:
            s1, s2 = int(l1 ** 0.5), int(l2 ** 0.5)
            resized = F.interpolate(pre.permute(1, 0).reshape(1, h1, s1, s1).float(), size=(s2, s2), mode="bicubic")
            sd[key] = resized.reshape(h2, l2).permute(1, 0).contiguous().to(pre.dtype)
    return sd


def _read(filename: str, map_location, allow_pickle: bool):
    """Local path, or an http(s):// URL through torch.hub (the reference's ``_load_checkpoint`` also resolves
    ``modelzoo://`` / ``open-mmlab://`` aliases through mmcv's model-zoo tables: not reproduced -- pass the URL itself).
    Tensors-only unpickling first; arbitrary-object unpickling (code execution on a hostile file) only on opt-in."""
    if filename.startswith(("http://", "https://")):
        return torch.hub.load_state_dict_from_url(filename, map_location=map_location, weights_only=not allow_pickle)
    if filename.startswith(("modelzoo://", "open-mmlab://", "torchvision://")):
        raise RuntimeError(f"{filename}: model-zoo aliases are not resolved here; pass a local path or an http(s) URL")
    try:
        return torch.load(filename, map_location=map_location, weights_only=True)
    except Exception as e:
        if not allow_pickle:
            raise RuntimeError(f"{filename} cannot be read with weights_only=True ({type(e).__name__}: {e}); if you trust the "
                               "file, call load_checkpoint(..., allow_pickle=True)") from e
        return torch.load(filename, map_location=map_location, weights_only=False)


def load_checkpoint(model: torch.nn.Module, filename: str, map_location="cpu", strict: bool = False, allow_pickle: bool = False):
    ckpt = _read(filename, map_location, allow_pickle)
    if not isinstance(ckpt, dict):
        raise RuntimeError(f"No state_dict found in checkpoint file {filename}")
    state = ckpt.get("state_dict", ckpt.get("model", ckpt))
    sd = adapt_state_dict(model, state)
    own = model.state_dict()
    usable = {k: v for k, v in sd.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}
    skipped = sorted(set(sd) - set(usable))
    missing = sorted(k for k in own if k not in usable and not k.endswith("relative_position_index"))
    if strict and (skipped or missing):
        raise RuntimeError(f"checkpoint mismatch: unexpected/mismatched {skipped[:5]}..., missing {missing[:5]}...")
    if skipped:
        _log.info("checkpoint keys not used by the backbone: %d (e.g. %s)", len(skipped), skipped[:3])
    if missing:
        _log.warning("backbone keys missing from the checkpoint: %s", missing[:8])
    model.load_state_dict(usable, strict=False)
    return ckpt
