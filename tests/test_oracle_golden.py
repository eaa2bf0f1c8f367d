"""The oracle (oracle/swin_oracle.py) against fixtures produced by the reference itself
(oracle/make_golden.py).  CPU only.  Bit-exact for index ops; float cases were generated in
fp64 and stored as fp32, so the restatement run in fp64 must agree to ~1e-7 rel-L2."""
import os

import numpy as np
import pytest
import torch

from oracle import swin_oracle as so
from oracle.make_golden import INDEX_CASES, TINY, rnd
from conftest import GOLDEN

TOL = 2e-6


def g(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.mark.parametrize("ci", range(len(INDEX_CASES)))
def test_index_ops_bit_exact(ci):
    d = g("index_ops.npz")
    B, H, W, C, ws = INDEX_CASES[ci]
    Hp, Wp = so.padded_hw(H, W, ws)
    x = np.random.default_rng(100 + ci).integers(-30000, 30000, size=(B, Hp, Wp, C)).astype(np.float32)
    part = so.window_partition_np(x, ws)
    assert np.array_equal(part.astype(np.int16), d[f"part{ci}"])
    assert np.array_equal(so.window_reverse_np(part, ws, Hp, Wp), x)
    xs = np.random.default_rng(200 + ci).integers(1, 30000, size=(B, H, W, C)).astype(np.float32)
    for shift in (0, ws // 2):
        got = so.shift_gather(torch.from_numpy(xs).reshape(B, H * W, C), H, W, ws, shift).numpy()
        assert np.array_equal(got.astype(np.int16), d[f"gather{ci}_s{shift}"])
        ww = np.random.default_rng(300 + ci + shift).integers(1, 30000, size=got.shape).astype(np.float32)
        back = so.shift_scatter(torch.from_numpy(ww), B, H, W, ws, shift).numpy()
        assert np.array_equal(back.astype(np.int16), d[f"scatter{ci}_s{shift}"])
    assert np.array_equal(so.shift_mask_np(H, W, ws, ws // 2), d[f"mask{ci}"])
    assert np.array_equal(so.relative_position_index_np(ws), d[f"relidx{ci}"])


def test_gather_scatter_roundtrip_property():
    # size-independent property: scatter(gather(x)) == x on every valid pixel, at the cfg2 stage-1 geometry
    H, W, ws = 100, 167, 7
    for shift in (0, 3):
        idx = so.gather_index(H, W, ws, shift)
        valid = idx[idx >= 0]
        assert valid.size == H * W and np.array_equal(np.sort(valid), np.arange(H * W))


@pytest.mark.parametrize("name,B_,nW", [("nomask", 5, 0), ("mask", 6, 3)])
def test_window_attention(name, B_, nW):
    d = g("window_attention.npz")
    C, nH, ws = 64, 2, 7
    shapes = {"relative_position_bias_table": ((2 * ws - 1) ** 2, nH), "qkv.weight": (3 * C, C), "qkv.bias": (3 * C,),
              "proj.weight": (C, C), "proj.bias": (C,)}
    p = {k: v.requires_grad_(True) for k, v in so.seeded_params(shapes, seed=11, dtype=torch.float64).items()}
    x = torch.from_numpy(rnd(21, (B_, ws * ws, C))).double().requires_grad_(True)
    cot = torch.from_numpy(rnd(22, (B_, ws * ws, C))).double()
    mask = torch.from_numpy(d[f"{name}_maskin"]).double() if nW else None
    y = so.window_attention(x, p, "", nH, ws, mask)
    (y * cot).sum().backward()
    assert so.rel_l2(y, torch.from_numpy(d[f"{name}_y"])) < TOL
    assert so.rel_l2(x.grad, torch.from_numpy(d[f"{name}_dx"])) < TOL
    for k, v in p.items():
        assert so.rel_l2(v.grad, torch.from_numpy(d[f"{name}_g_{k}"])) < TOL, k


@pytest.mark.parametrize("shift", [0, 3])
def test_swin_block(shift):
    d = g("swin_block.npz")
    C, nH, ws, B, H, W = 64, 2, 7, 2, 10, 13
    hid = 4 * C
    shapes = {"norm1.weight": (C,), "norm1.bias": (C,),
              "attn.relative_position_bias_table": ((2 * ws - 1) ** 2, nH),
              "attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
              "attn.proj.weight": (C, C), "attn.proj.bias": (C,),
              "norm2.weight": (C,), "norm2.bias": (C,),
              "mlp.fc1.weight": (hid, C), "mlp.fc1.bias": (hid,), "mlp.fc2.weight": (C, hid), "mlp.fc2.bias": (C,)}
    p = {k: v.requires_grad_(True) for k, v in so.seeded_params(shapes, seed=31, dtype=torch.float64).items()}
    x = torch.from_numpy(rnd(41, (B, H * W, C))).double().requires_grad_(True)
    cot = torch.from_numpy(rnd(42, (B, H * W, C))).double()
    y = so.swin_block(x, H, W, p, "", nH, ws, shift)
    (y * cot).sum().backward()
    assert so.rel_l2(y, torch.from_numpy(d[f"s{shift}_y"])) < TOL
    assert so.rel_l2(x.grad, torch.from_numpy(d[f"s{shift}_dx"])) < TOL
    for k, v in p.items():
        assert so.rel_l2(v.grad, torch.from_numpy(d[f"s{shift}_g_{k}"])) < TOL, k


def test_patch_merging():
    d = g("patch_merging.npz")
    C, B, H, W = 32, 2, 5, 9
    shapes = {"reduction.weight": (2 * C, 4 * C), "norm.weight": (4 * C,), "norm.bias": (4 * C,)}
    p = {k: v.requires_grad_(True) for k, v in so.seeded_params(shapes, seed=51, dtype=torch.float64).items()}
    x = torch.from_numpy(rnd(61, (B, H * W, C))).double().requires_grad_(True)
    y = so.patch_merging(x, H, W, p, "")
    cot = torch.from_numpy(rnd(62, tuple(y.shape))).double()
    (y * cot).sum().backward()
    assert so.rel_l2(y, torch.from_numpy(d["y"])) < TOL
    assert so.rel_l2(x.grad, torch.from_numpy(d["dx"])) < TOL
    for k, v in p.items():
        assert so.rel_l2(v.grad, torch.from_numpy(d["g_" + k])) < TOL, k


def test_backbone_tiny():
    d = g("backbone_tiny.npz")
    shapes = so.param_shapes(TINY["embed_dim"], TINY["depths"], TINY["num_heads"], TINY["window_size"],
                             out_indices=TINY["out_indices"])
    p = {k: v.requires_grad_(True) for k, v in so.seeded_params(shapes, seed=71, dtype=torch.float64).items()}
    img = torch.from_numpy(rnd(81, (2, 3, 50, 70))).double().requires_grad_(True)
    outs = so.backbone_forward(img, p, **TINY)
    loss = 0
    for i, o in enumerate(outs):
        assert so.rel_l2(o, torch.from_numpy(d[f"out{i}"])) < TOL
        loss = loss + (o * torch.from_numpy(rnd(90 + i, tuple(o.shape))).double()).sum()
    loss.backward()
    assert so.rel_l2(img.grad, torch.from_numpy(d["dimg"])) < TOL
    for k, v in p.items():
        assert so.rel_l2(v.grad, torch.from_numpy(d["g_" + k])) < 5e-6, k


def test_param_shapes_match_swin_t_state_dict():
    d = g("backbone_tiny.npz")
    shapes = so.param_shapes(96, (2, 2, 6, 2), (3, 6, 12, 24))
    ref = {k: tuple(int(t) for t in s.split(",")) for k, s in zip(d["swin_t_keys"], d["swin_t_shapes"])}
    ref = {k: v for k, v in ref.items() if not k.endswith("relative_position_index")}
    assert ref == shapes
    assert list(ref.keys()) == list(shapes.keys())      # module order too
