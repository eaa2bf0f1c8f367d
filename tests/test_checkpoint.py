"""Host logic of row f4: the checkpoint loader reproduces the reference loader's Swin fix-ups
(mmcv_custom/checkpoint.py:286-356).  CPU only."""
import os

import pytest
import torch
import torch.nn.functional as F


def _net(**kw):
    import swin_b200
    cfg = dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=7, out_indices=(0, 1))
    cfg.update(kw)
    return swin_b200.SwinTransformer(**cfg)


def test_roundtrip_plain_model_and_module_prefix(tmp_path):
    src = _net()
    src.init_weights()
    for wrap in (lambda sd: sd, lambda sd: {"state_dict": sd}, lambda sd: {"model": {"module." + k: v for k, v in sd.items()}}):
        f = os.path.join(tmp_path, "c.pth")
        torch.save(wrap(src.state_dict()), f)
        dst = _net()
        dst.init_weights(pretrained=f)
        for k, v in src.state_dict().items():
            assert torch.equal(dst.state_dict()[k], v), k


def test_moby_encoder_prefix_and_head_keys_ignored(tmp_path):
    src = _net()
    src.init_weights()
    sd = {"encoder." + k: v for k, v in src.state_dict().items()}
    sd["projector.weight"] = torch.zeros(3, 3)
    f = os.path.join(tmp_path, "m.pth")
    torch.save({"model": sd}, f)
    dst = _net()
    dst.init_weights(pretrained=f)
    assert torch.equal(dst.layers[1].blocks[1].mlp.fc2.weight, src.layers[1].blocks[1].mlp.fc2.weight)


def test_relative_position_bias_table_is_bicubically_resized(tmp_path):
    src = _net(window_size=7)
    for p in src.parameters():
        torch.nn.init.normal_(p, std=0.3)
    f = os.path.join(tmp_path, "w7.pth")
    torch.save(src.state_dict(), f)
    dst = _net(window_size=12)
    dst.init_weights(pretrained=f)
    t7 = src.layers[0].blocks[0].attn.relative_position_bias_table.detach()
    want = F.interpolate(t7.permute(1, 0).view(1, 1, 13, 13), size=(23, 23), mode="bicubic").view(1, 529).permute(1, 0)
    assert torch.allclose(dst.layers[0].blocks[0].attn.relative_position_bias_table, want)
    assert torch.equal(dst.layers[0].blocks[0].attn.qkv.weight, src.layers[0].blocks[0].attn.qkv.weight)


def test_absolute_pos_embed_reshape(tmp_path):
    src = _net(ape=True, pretrain_img_size=56)
    sd = src.state_dict()
    n, c, h, w = sd["absolute_pos_embed"].shape
    flat = torch.randn(n, h * w, c)
    sd["absolute_pos_embed"] = flat
    f = os.path.join(tmp_path, "a.pth")
    torch.save(sd, f)
    dst = _net(ape=True, pretrain_img_size=56)
    dst.init_weights(pretrained=f)
    assert torch.equal(dst.absolute_pos_embed, flat.view(n, h, w, c).permute(0, 3, 1, 2))
    with pytest.raises(TypeError):
        dst.init_weights(pretrained=5)
