"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the authoring container only).

    python oracle/make_golden.py

Inputs and weights come from numpy Generators with fixed seeds (re-creatable on any box; a
checksum of each is stored to detect generator drift); outputs / gradients are what the
reference module at /root/reference produces on CPU in fp32 (fp64 for the float cases, then
stored as fp32, so the fixtures sit at the reference's own noise floor).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader, swin_oracle as so  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def rnd(seed, shape, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


def chk(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), (a * a).sum()], dtype=np.float64)


INDEX_CASES = [  # (B, H, W, C, ws)
    (2, 10, 13, 8, 7),
    (1, 7, 7, 8, 7),
    (1, 20, 33, 8, 12),
    (3, 5, 9, 16, 7),
    (1, 14, 14, 8, 7),
]


def gen_index(ref):
    d = {}
    for ci, (B, H, W, C, ws) in enumerate(INDEX_CASES):
        Hp, Wp = so.padded_hw(H, W, ws)
        s = ws // 2
        x = np.random.default_rng(100 + ci).integers(-30000, 30000, size=(B, Hp, Wp, C)).astype(np.float32)
        xt = torch.from_numpy(x)
        part = ref.window_partition(xt, ws)
        d[f"part{ci}"] = part.numpy().astype(np.int16)
        back = ref.window_reverse(part, ws, Hp, Wp)
        assert torch.equal(back, xt)
        # composite a3: pad -> roll -> partition exactly as REF:214-231 does it
        xs = np.random.default_rng(200 + ci).integers(1, 30000, size=(B, H, W, C)).astype(np.float32)
        t = torch.from_numpy(xs)
        for shift in (0, s):
            p = torch.nn.functional.pad(t, (0, 0, 0, Wp - W, 0, Hp - H))
            if shift > 0:
                p = torch.roll(p, shifts=(-shift, -shift), dims=(1, 2))
            w = ref.window_partition(p, ws).view(-1, ws * ws, C)
            d[f"gather{ci}_s{shift}"] = w.numpy().astype(np.int16)
            # a4 reverse composite REF:236-249 applied to a fresh random window tensor
            ww = np.random.default_rng(300 + ci + shift).integers(1, 30000, size=tuple(w.shape)).astype(np.float32)
            r = ref.window_reverse(torch.from_numpy(ww).view(-1, ws, ws, C), ws, Hp, Wp)
            if shift > 0:
                r = torch.roll(r, shifts=(shift, shift), dims=(1, 2))
            r = r[:, :H, :W, :].contiguous().view(B, H * W, C)
            d[f"scatter{ci}_s{shift}"] = r.numpy().astype(np.int16)
        # a5 mask, via BasicLayer.forward's own code path: capture what the blocks receive
        layer = ref.BasicLayer(dim=32, depth=2, num_heads=1, window_size=ws, drop_path=[0.0, 0.0])
        seen = {}

        def hook(mod, args):
            seen["mask"] = args[1].detach().clone()
        h = layer.blocks[1].register_forward_pre_hook(hook)
        with torch.no_grad():
            layer(torch.zeros(1, H * W, 32), H, W)
        h.remove()
        d[f"mask{ci}"] = seen["mask"].numpy().astype(np.float32)
        att = ref.WindowAttention(32, (ws, ws), 1)
        d[f"relidx{ci}"] = att.relative_position_index.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "index_ops.npz"), **d)


def load_params(module, params, prefix=""):
    sd = module.state_dict()
    for k in list(sd.keys()):
        if k.endswith("relative_position_index"):
            continue
        sd[k] = params[prefix + k].to(sd[k].dtype)
    module.load_state_dict(sd)


def grads_of(module, prefix=""):
    return {prefix + k: v.grad.detach().float().numpy() for k, v in module.named_parameters()}


def gen_attention(ref):
    d = {}
    C, nH, ws = 64, 2, 7
    N = ws * ws
    shapes = {"relative_position_bias_table": ((2 * ws - 1) ** 2, nH), "qkv.weight": (3 * C, C), "qkv.bias": (3 * C,),
              "proj.weight": (C, C), "proj.bias": (C,)}
    params = so.seeded_params(shapes, seed=11, dtype=torch.float64)
    for name, B_, nW in (("nomask", 5, 0), ("mask", 6, 3)):
        m = ref.WindowAttention(C, (ws, ws), nH).double()
        load_params(m, params)
        x = torch.from_numpy(rnd(21, (B_, N, C))).double().requires_grad_(True)
        cot = torch.from_numpy(rnd(22, (B_, N, C))).double()
        mask = None
        if nW:
            full = so.shift_mask_np(10, 13, ws, 3)                       # (4,49,49); take the 3 non-trivial
            mask = torch.from_numpy(full[1:4]).double()
            d[f"{name}_maskin"] = mask.float().numpy()
        y = m(x, mask)
        (y * cot).sum().backward()
        d[f"{name}_y"] = y.detach().float().numpy()
        d[f"{name}_dx"] = x.grad.float().numpy()
        for k, v in grads_of(m).items():
            d[f"{name}_g_{k}"] = v
        d[f"{name}_chk"] = np.concatenate([chk(x.detach()), chk(cot)])
    np.savez_compressed(os.path.join(OUT, "window_attention.npz"), **d)


def gen_block(ref):
    d = {}
    C, nH, ws, B, H, W = 64, 2, 7, 2, 10, 13
    hid = 4 * C
    shapes = {"norm1.weight": (C,), "norm1.bias": (C,),
              "attn.relative_position_bias_table": ((2 * ws - 1) ** 2, nH),
              "attn.qkv.weight": (3 * C, C), "attn.qkv.bias": (3 * C,),
              "attn.proj.weight": (C, C), "attn.proj.bias": (C,),
              "norm2.weight": (C,), "norm2.bias": (C,),
              "mlp.fc1.weight": (hid, C), "mlp.fc1.bias": (hid,), "mlp.fc2.weight": (C, hid), "mlp.fc2.bias": (C,)}
    params = so.seeded_params(shapes, seed=31, dtype=torch.float64)
    for shift in (0, 3):
        layer = ref.BasicLayer(dim=C, depth=2, num_heads=nH, window_size=ws, drop_path=[0.0, 0.0]).double()
        blk = layer.blocks[1 if shift else 0]
        load_params(blk, params)
        # obtain the canonical mask from the layer's own code
        seen = {}
        h = layer.blocks[1].register_forward_pre_hook(lambda mod, a: seen.__setitem__("mask", a[1].detach().clone()))
        with torch.no_grad():
            layer(torch.zeros(1, H * W, C, dtype=torch.float64), H, W)
        h.remove()
        x = torch.from_numpy(rnd(41, (B, H * W, C))).double().requires_grad_(True)
        cot = torch.from_numpy(rnd(42, (B, H * W, C))).double()
        blk.H, blk.W = H, W
        y = blk(x, seen["mask"].double())
        (y * cot).sum().backward()
        d[f"s{shift}_y"] = y.detach().float().numpy()
        d[f"s{shift}_dx"] = x.grad.float().numpy()
        for k, v in grads_of(blk).items():
            d[f"s{shift}_g_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "swin_block.npz"), **d)


def gen_merge(ref):
    d = {}
    C, B, H, W = 32, 2, 5, 9
    shapes = {"reduction.weight": (2 * C, 4 * C), "norm.weight": (4 * C,), "norm.bias": (4 * C,)}
    params = so.seeded_params(shapes, seed=51, dtype=torch.float64)
    m = ref.PatchMerging(C).double()
    load_params(m, params)
    x = torch.from_numpy(rnd(61, (B, H * W, C))).double().requires_grad_(True)
    y = m(x, H, W)
    cot = torch.from_numpy(rnd(62, tuple(y.shape))).double()
    (y * cot).sum().backward()
    d["y"] = y.detach().float().numpy()
    d["dx"] = x.grad.float().numpy()
    for k, v in grads_of(m).items():
        d["g_" + k] = v
    np.savez_compressed(os.path.join(OUT, "patch_merging.npz"), **d)


TINY = dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=7, out_indices=(0, 1))


def gen_backbone(ref):
    d = {}
    shapes = so.param_shapes(TINY["embed_dim"], TINY["depths"], TINY["num_heads"], TINY["window_size"],
                             out_indices=TINY["out_indices"])
    params = so.seeded_params(shapes, seed=71, dtype=torch.float64)
    net = ref.SwinTransformer(drop_path_rate=0.0, **TINY).double()
    assert set(k for k in net.state_dict() if not k.endswith("relative_position_index")) == set(shapes), \
        "oracle.param_shapes disagrees with the reference state_dict"
    load_params(net, params)
    net.train()
    img = torch.from_numpy(rnd(81, (2, 3, 50, 70))).double().requires_grad_(True)
    outs = net(img)
    loss = 0
    for i, o in enumerate(outs):
        cot = torch.from_numpy(rnd(90 + i, tuple(o.shape))).double()
        loss = loss + (o * cot).sum()
        d[f"out{i}"] = o.detach().float().numpy()
    loss.backward()
    d["dimg"] = img.grad.float().numpy()
    for k, v in grads_of(net).items():
        d["g_" + k] = v
    d["chk_img"] = chk(img.detach())
    d["state_keys"] = np.array(sorted(net.state_dict().keys()))
    # Swin-T state_dict contract (names + shapes) for the boundary test
    full = ref.SwinTransformer()
    d["swin_t_keys"] = np.array(list(full.state_dict().keys()))
    d["swin_t_shapes"] = np.array([",".join(map(str, v.shape)) for v in full.state_dict().values()])
    np.savez_compressed(os.path.join(OUT, "backbone_tiny.npz"), **d)


def main():
    assert ref_loader.available(), "reference not mounted"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    ref = ref_loader.load()
    gen_index(ref)
    gen_attention(ref)
    gen_block(ref)
    gen_merge(ref)
    gen_backbone(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
