"""Live check of the CPU oracle against the UNMODIFIED reference module (mmdet/models/backbones/swin_transformer.py),
imported from /root/reference through the stub loader.  Skipped where the reference tree is absent (the GPU box): there
the committed fixtures of tests/golden (generated from this same import by oracle/make_golden.py) carry the pinning."""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import swin_oracle as so

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not present")


def _load_ref(cfg, params):
    ref = ref_loader.load()
    net = ref.SwinTransformer(drop_path_rate=0.0, ape="absolute_pos_embed" in params, pretrain_img_size=56, **cfg)
    sd = net.state_dict()
    for k in sd:
        if k.endswith("relative_position_index") or k.endswith("attn_mask"):
            continue
        sd[k] = params[k].clone()
    net.load_state_dict(sd)
    net.train()           # the reference's train() returns None (REF:627-630)
    return net


@pytest.mark.parametrize("cfg,B,HW", [
    (dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=7, out_indices=(0, 1)), 2, (50, 70)),
    (dict(embed_dim=96, depths=[2, 2, 2, 2], num_heads=[3, 6, 12, 24], window_size=7, out_indices=(0, 1, 2, 3)), 1, (224, 300)),
    (dict(embed_dim=32, depths=[2], num_heads=[1], window_size=12, out_indices=(0,)), 1, (90, 100)),
    (dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=7, out_indices=(0, 1), _ape=True), 2, (60, 84)),
])
def test_oracle_matches_live_reference_forward_and_backward(cfg, B, HW):
    torch.manual_seed(0)
    cfg = dict(cfg)
    ape = cfg.pop("_ape", False)
    shapes = so.param_shapes(cfg["embed_dim"], cfg["depths"], cfg["num_heads"], cfg["window_size"], out_indices=cfg["out_indices"])
    if ape:                                              # REF:513-517: (1, C, pretrain/patch, pretrain/patch)
        shapes["absolute_pos_embed"] = (1, cfg["embed_dim"], 14, 14)
    params = so.seeded_params(shapes, seed=3)
    img = torch.from_numpy(np.random.default_rng(1).standard_normal((B, 3) + HW).astype(np.float32))
    net = _load_ref(cfg, params)
    im_r = img.clone().requires_grad_(True)
    outs_r = net(im_r)
    cots = [torch.from_numpy(np.random.default_rng(10 + i).standard_normal(tuple(o.shape)).astype(np.float32)) for i, o in enumerate(outs_r)]
    sum((o * c).sum() for o, c in zip(outs_r, cots)).backward()
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    im_o = img.clone().requires_grad_(True)
    outs_o = so.backbone_forward(im_o, p, **cfg)
    sum((o * c).sum() for o, c in zip(outs_o, cots)).backward()
    for i, (a, b) in enumerate(zip(outs_o, outs_r)):
        assert a.shape == b.shape
        assert so.rel_l2(a, b) < 5e-6, (i, so.rel_l2(a, b))
    assert so.rel_l2(im_o.grad, im_r.grad) < 1e-5
    worst = max((so.rel_l2(p[k].grad, v.grad), k) for k, v in net.named_parameters())
    assert worst[0] < 2e-5, worst


def test_state_dict_keys_match_reference():
    ref = ref_loader.load()
    net = ref.SwinTransformer()
    want = {k: tuple(v.shape) for k, v in net.state_dict().items() if not k.endswith("relative_position_index")}
    got = so.param_shapes(96, [2, 2, 6, 2], [3, 6, 12, 24], 7)
    assert want == {k: tuple(v) for k, v in got.items()}
