"""Round-2 additions on the GPU: stand-alone Mlp, frozen stages, torch-format optimizer state, the benchmarked B=16 shape
checked on device against the oracle-pinned fp32 kernels, and the 2-GPU NCCL data-parallel equivalence."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import swin_oracle as so
from oracle.make_golden import TINY, rnd

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_mlp_standalone_matches_oracle(mode, tol):
    """Mlp.forward on its own (REF:32-38), forward and all gradients."""
    import swin_b200
    C, hid = 64, 256
    shapes = {"fc1.weight": (hid, C), "fc1.bias": (hid,), "fc2.weight": (C, hid), "fc2.bias": (C,)}
    params = so.seeded_params(shapes, seed=13)
    m = swin_b200.swin_transformer.Mlp(C, hid, compute_dtype=mode)
    m.load_state_dict({k: v.float() for k, v in params.items()})
    m = m.to(DEV)
    x = torch.from_numpy(rnd(5, (3, 37, C)))
    cot = torch.from_numpy(rnd(6, (3, 37, C)))
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xr = x.clone().requires_grad_(True)
    (so.mlp(xr, p, "") * cot).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    y = m(xg)
    (y * cot.to(DEV)).sum().backward()
    assert so.rel_l2(y, so.mlp(x, params, "")) < tol
    assert so.rel_l2(xg.grad, xr.grad) < tol
    for k, v in m.named_parameters():
        assert so.rel_l2(v.grad, p[k].grad) < tol, k


def test_frozen_stages_skip_weight_gradients_and_match():
    """frozen_stages=2 (REF:557-572): patch_embed and stage 0 take no gradients, no weight-gradient kernels are launched for
    them, and the trainable part's gradients equal the unfrozen model's."""
    import swin_b200
    from swin_b200 import ops
    shapes = so.param_shapes(TINY["embed_dim"], TINY["depths"], TINY["num_heads"], TINY["window_size"], out_indices=TINY["out_indices"])
    params = so.seeded_params(shapes, seed=71)
    img = torch.from_numpy(rnd(81, (2, 3, 50, 70))).to(DEV)
    res, launches = {}, {}
    for fs in (-1, 2):
        net = swin_b200.SwinTransformer(drop_path_rate=0.0, compute_dtype="fp32", frozen_stages=fs, **TINY)
        sd = net.state_dict()
        for k in sd:
            if not k.endswith("relative_position_index"):
                sd[k] = params[k].float()
        net.load_state_dict(sd)
        net = net.to(DEV).train()
        l0 = ops.LAUNCHES
        sum(o.sum() for o in net(img)).backward()
        launches[fs] = ops.LAUNCHES - l0
        res[fs] = {k: (None if v.grad is None else v.grad.clone()) for k, v in net.named_parameters()}
    assert launches[2] < launches[-1]
    for k, g in res[2].items():
        if k.startswith("patch_embed") or k.startswith("layers.0."):
            assert g is None, k
        else:
            assert so.rel_l2(g, res[-1][k]) < 1e-5, k


def test_fused_adamw_is_a_torch_optimizer_with_torch_state():
    """param_groups drive lr (schedulers / mmcv LrUpdaterHook), state_dict is torch.optim.AdamW's layout both ways."""
    import swin_b200
    from swin_b200.optim import FusedAdamW
    torch.manual_seed(0)
    kw = dict(embed_dim=32, depths=[2], num_heads=[1], out_indices=(0,), drop_path_rate=0.0, compute_dtype="fp32")
    net = swin_b200.SwinTransformer(**kw).to(DEV).train()
    twin = swin_b200.SwinTransformer(**kw).to(DEV).train()
    twin.load_state_dict(net.state_dict())
    opt = FusedAdamW(net, lr=1e-3, weight_decay=0.05)
    assert isinstance(opt, torch.optim.Optimizer) and all("lr" in g for g in opt.param_groups)
    name_of = {id(p): n for n, p in twin.named_parameters()}
    groups = [{"params": [p], "weight_decay": 0.0 if ("norm" in name_of[id(p)] or "relative_position_bias_table" in name_of[id(p)]) else 0.05}
              for p in twin.parameters()]
    ref = torch.optim.AdamW(groups, lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)
    sched_ref = torch.optim.lr_scheduler.StepLR(ref, step_size=1, gamma=0.5)
    x = torch.randn(2, 3, 56, 56, device=DEV)
    for it in range(3):
        for n_, o_ in ((net, opt), (twin, ref)):
            o_.zero_grad()
            sum(o.square().mean() for o in n_(x)).backward()
        # identical gradients into both optimizers (the kernels are deterministic up to atomics: copy to be exact)
        for (k, p), (_, q) in zip(net.named_parameters(), twin.named_parameters()):
            q.grad = p.grad.clone()
        opt.step(); ref.step()
        sched.step(); sched_ref.step()
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-3 * 0.5 ** 3)
    for (k, p), (_, q) in zip(net.named_parameters(), twin.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), k
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    # torch.optim.AdamW reads it ...
    flat = torch.optim.AdamW([{"params": g["params"]} for g in opt.param_groups], lr=1.0)
    flat.load_state_dict(sd)
    assert float(flat.state[opt.param_groups[0]["params"][0]]["step"]) == 3.0
    # ... and a fresh FusedAdamW resumes from it
    again = FusedAdamW(net, lr=1e-3, weight_decay=0.05)
    again.load_state_dict(sd)
    p0 = again.param_groups[0]["params"][0]
    assert torch.equal(again.state[p0]["exp_avg"], opt.state[opt.param_groups[0]["params"][0]]["exp_avg"])


def test_benchmarked_shape_b16_bf16_vs_fp32_kernels_on_device():
    """The benchmarked configuration itself (Swin-T, B=16, 3x800x1333: its split-K factors, persistent-grid walks and tile
    choices) -- bf16 tcgen05 path against the fp32 FFMA kernels (which are pinned to the oracle at <=1e-4) on the same
    seeded weights: outputs, input gradient and every parameter gradient.  The oracle itself is too slow at B=16."""
    import swin_b200
    cfg = dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7)
    shapes = so.param_shapes(**cfg)
    params = so.seeded_params(shapes, seed=7)
    img = torch.from_numpy(np.random.default_rng(1).standard_normal((16, 3, 800, 1333)).astype(np.float32)).to(DEV)
    res = {}
    cots = None
    for mode in ("fp32", "bf16"):
        net = swin_b200.SwinTransformer(drop_path_rate=0.0, compute_dtype=mode, **cfg)
        sd = net.state_dict()
        for k in sd:
            if not k.endswith("relative_position_index"):
                sd[k] = params[k].float()
        net.load_state_dict(sd)
        net = net.to(DEV).train()
        im = img.clone().requires_grad_(True)
        outs = net(im)
        if cots is None:
            g = torch.Generator(device=DEV).manual_seed(5)
            cots = [torch.randn(o.shape, device=DEV, generator=g) for o in outs]
        torch.autograd.backward(outs, cots)
        torch.cuda.synchronize()
        res[mode] = ([o.detach().clone() for o in outs], im.grad.clone(), {k: v.grad.clone() for k, v in net.named_parameters()})
        del net, outs, im
        torch.cuda.empty_cache()
    for i, (a, b) in enumerate(zip(res["bf16"][0], res["fp32"][0])):
        assert so.rel_l2(a, b) < 2e-2, (i, so.rel_l2(a, b))
    assert so.rel_l2(res["bf16"][1], res["fp32"][1]) < 2e-2
    path = os.path.join(GOLDEN, "bf16_exceptions.json")
    ceil = {}
    if os.path.isfile(path):
        with open(path) as f:
            ceil = json.load(f).get("swin_t_B16_800x1333_vs_fp32_kernels", {})
    rows = sorted(((so.rel_l2(res["bf16"][2][k], v), k) for k, v in res["fp32"][2].items()), reverse=True)
    over = [(k, e) for e, k in rows if e >= 2e-2]
    print(f"[parity] B=16 800x1333 bf16 vs fp32 kernels: worst parameter gradients {rows[:5]}; {len(over)} above 2e-2")
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_b16_vs_fp32.json"), "w") as f:
            json.dump([{"tensor": k, "rel_l2": e} for e, k in rows], f, indent=1)
    except OSError:
        pass
    bad = [(k, e) for k, e in over if e > ceil.get(k, 0.0)]
    assert not bad, bad


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
def test_two_gpu_nccl_sharded_gradients_equal_full_batch():
    """the real backbone, BucketedGradAllReduce over NCCL on 2 GPUs, per-rank half batches -- the averaged
    gradients equal the single-GPU full-batch gradients (fp32 kernels, tolerance 1e-5)."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(ROOT, "tests", "ddp_nccl_worker.py")], capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert "DDP_NCCL_OK" in r.stdout


@pytest.mark.parametrize("nH,B_,mask_kind", [(3, 22, "none"), (3, 24, "canon"), (3, 24, "tensor"), (2, 7, "none"), (1, 12, "canon"),
                                             (3, 1, "none"), (3, 4000, "canon"),
                                             # streamed-weight mode: C = 128 / 192 (two X slots), 256 / 384 (one X slot)
                                             (4, 24, "canon"), (4, 7, "none"), (6, 24, "tensor"), (6, 4000, "canon"), (8, 24, "canon"),
                                             (12, 24, "canon"), (12, 1, "none"), (12, 4000, "canon"), (12, 777, "none")])
def test_fused_qkv_attention_matches_separate_kernels(nH, B_, mask_kind):
    """swin_window_attn_qkv_fwd (qkv projection inside the attention kernel, REF:128-150) against the qkv GEMM + the
    stand-alone attention kernel on the same bf16 inputs, and both against the fp32 arithmetic of the oracle."""
    from swin_b200 import ops
    from swin_b200 import _lib as L
    ws, N, C = 7, 49, 32 * nH
    if not ops.window_attn_qkv_supported(C, nH, ws):
        pytest.skip("shape not supported by the fused kernel")
    g = torch.Generator(device=DEV).manual_seed(nH * 1000 + B_)
    xw = (torch.randn(B_ * N, C, device=DEV, generator=g)).bfloat16()
    w = (torch.randn(3 * C, C, device=DEV, generator=g) / C ** 0.5).bfloat16()
    b = torch.randn(3 * C, device=DEV, generator=g) * 0.3
    table = torch.randn((2 * ws - 1) ** 2, nH, device=DEV, generator=g) * 0.5
    bias = ops.rel_bias_expand(table, ws)
    mask = mnz = None
    canon = (0, 0)
    if mask_kind != "none":
        grid = {24: (2, 2), 12: (2, 2), 4000: (4, 5)}[B_]
        mask = torch.from_numpy(so.shift_mask_np(grid[0] * ws, grid[1] * ws, ws, 3)).to(DEV)
        mnz = ops.mask_nonzero(mask)
        if mask_kind == "canon":
            canon = grid
        else:
            mask = mask * (torch.rand(mask.shape, device=DEV, generator=g) > 0.3).float()      # an arbitrary additive mask
            mnz = ops.mask_nonzero(mask)
    scale = 32 ** -0.5
    qkv_ref = ops.gemm(xw, w, B_ * N, 3 * C, C, bias=b)
    o_ref, lse_ref = ops.window_attn_fwd(qkv_ref.view(B_, N, 3 * C), bias, mask, B_, nH, ws, scale, mnz, canon)
    o, lse, qkv = ops.window_attn_qkv_fwd(xw, w, b, bias, mask, B_, nH, ws, scale, mnz, canon, want_qkv=True)
    torch.cuda.synchronize()
    assert so.rel_l2(qkv.view(-1, 3 * C), qkv_ref) < 1e-3
    assert so.rel_l2(o, o_ref) < 4e-3, so.rel_l2(o, o_ref)
    assert torch.allclose(lse, lse_ref, rtol=1e-3, atol=2e-3)
    # inference variant: no qkv, no lse
    o2, lse2, qkv2 = ops.window_attn_qkv_fwd(xw, w, b, bias, mask, B_, nH, ws, scale, mnz, canon, want_qkv=False, want_lse=False)
    assert qkv2 is None and lse2 is None and torch.equal(o2, o)
    # fp32 arithmetic
    xf, wf = xw.float().cpu(), w.float().cpu()
    qkvf = (xf @ wf.t() + b.cpu()).view(B_, N, 3, nH, 32).permute(2, 0, 3, 1, 4)
    s = (qkvf[0] * scale) @ qkvf[1].transpose(-1, -2) + bias.cpu()[None]
    if mask is not None:
        s = s + mask.cpu()[torch.arange(B_) % mask.shape[0]][:, None]
    want = (torch.softmax(s, -1) @ qkvf[2]).transpose(1, 2).reshape(B_, N, C)
    assert so.rel_l2(o, want) < 1.5e-2, so.rel_l2(o, want)


def test_no_grad_forward_uses_fused_kernel_and_matches_training_forward():
    """Inference (torch.no_grad) runs stage-0 blocks through the fused QKV + attention kernel; outputs agree with the
    training-mode forward (two-kernel chain) on the same weights."""
    import swin_b200
    from swin_b200 import ops

    class Rec:
        def __init__(self): self.kinds = []
        def begin(self, kind, flops=0.0, nbytes=0.0): self.kinds.append(kind)
        def end(self): pass
    torch.manual_seed(1)
    net = swin_b200.SwinTransformer(embed_dim=96, depths=[2, 2], num_heads=[3, 6], out_indices=(0, 1), drop_path_rate=0.0).to(DEV).eval()
    x = torch.randn(2, 3, 120, 200, device=DEV)
    rec = Rec()
    ops.set_kernel_timer(rec)
    try:
        with torch.no_grad():
            outs_ng = net(x)
        n_fused = sum(k.startswith("attn_qkv_fwd") for k in rec.kinds)
        rec.kinds.clear()
        outs_g = net(x.clone().requires_grad_(True))
        n_fused_g = sum(k.startswith("attn_qkv_fwd") for k in rec.kinds)
    finally:
        ops.set_kernel_timer(None)
    assert n_fused == 2 and n_fused_g == 0          # the two stage-0 blocks (C = 96, resident weights); C = 192 keeps the two-kernel chain
    for a, b in zip(outs_ng, outs_g):
        assert so.rel_l2(a, b) < 5e-3


def test_window12_model_runs_in_bf16_mode_and_matches_fp32_mode():
    """Window 12 (the 384-pixel Swin-B/L configs): bf16 mode keeps every GEMM on tcgen05 and runs the attention core on the
    fp32-arithmetic kernels with bf16 storage; outputs and gradients agree with the fp32 mode (itself oracle-checked for window 12 at
    kernel level) within the bf16 tolerance.  Seeded noisy weights and random cotangents (with gamma = 1 and a plain sum as the
    loss, the gradient through the output LayerNorm is identically zero and every comparison would be noise against noise)."""
    import swin_b200
    cfg = dict(embed_dim=64, depths=[2, 2], num_heads=[2, 4], window_size=12, out_indices=(0, 1))
    shapes = so.param_shapes(cfg["embed_dim"], cfg["depths"], cfg["num_heads"], cfg["window_size"], out_indices=cfg["out_indices"])
    params = so.seeded_params(shapes, seed=21)
    nets = {}
    for mode in ("fp32", "bf16"):
        net = swin_b200.SwinTransformer(drop_path_rate=0.0, compute_dtype=mode, **cfg)
        sd = net.state_dict()
        for k in sd:
            if not k.endswith("relative_position_index"):
                sd[k] = params[k]
        net.load_state_dict(sd)
        nets[mode] = net.to(DEV).train()
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(2, 3, 150, 200, device=DEV, generator=g)
    outs_r = nets["fp32"](x)
    cots = [torch.randn(o.shape, device=DEV, generator=g) for o in outs_r]
    torch.autograd.backward(outs_r, cots)
    outs = nets["bf16"](x)
    torch.autograd.backward(outs, cots)
    torch.cuda.synchronize()
    for a, b in zip(outs, outs_r):
        assert so.rel_l2(a, b) < 2e-2
    errs = sorted(((so.rel_l2(p.grad, q.grad), k) for (k, p), (_, q) in zip(nets["bf16"].named_parameters(), nets["fp32"].named_parameters())), reverse=True)
    assert errs[0][0] < 4e-2, errs[:5]


def test_absolute_position_embedding_path_vs_oracle():
    """ape=True (REF:513-517, :604-607): the bicubic-interpolated absolute position embedding is added to the patch tokens before
    the first block; fp32 mode against the oracle (pinned to the live reference for this path by
    tests/test_oracle_vs_reference.py), outputs and every gradient incl. absolute_pos_embed."""
    import swin_b200
    cfg = dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=7, out_indices=(0, 1))
    shapes = so.param_shapes(cfg["embed_dim"], cfg["depths"], cfg["num_heads"], cfg["window_size"], out_indices=cfg["out_indices"])
    shapes["absolute_pos_embed"] = (1, 32, 14, 14)
    params = so.seeded_params(shapes, seed=11)
    img = torch.from_numpy(np.random.default_rng(2).standard_normal((2, 3, 60, 84)).astype(np.float32))
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    outs_r = so.backbone_forward(img, p, **cfg)
    sum(o.sum() for o in outs_r).backward()
    net = swin_b200.SwinTransformer(drop_path_rate=0.0, compute_dtype="fp32", ape=True, pretrain_img_size=56, **cfg)
    sd = net.state_dict()
    for k in sd:
        if not k.endswith("relative_position_index"):
            sd[k] = params[k]
    net.load_state_dict(sd)
    net = net.to(DEV).train()
    outs = net(img.to(DEV))
    sum(o.sum() for o in outs).backward()
    torch.cuda.synchronize()
    for a, b in zip(outs, outs_r):
        assert so.rel_l2(a, b) < 1e-4
    for k, v in net.named_parameters():
        assert so.rel_l2(v.grad, p[k].grad) < 1e-4, k
