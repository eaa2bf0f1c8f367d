"""The part of BASELINE.json configs[2] (Mask R-CNN Swin-T training step) that lies on this repo's path, as one measured step:

    backbone forward -> norm{i} + FPN lateral 1x1 convs on the token-major stage outputs (row f2, fpn.py)
    -> backward of both -> bucketed gradient all-reduce (N > 1) -> fused AdamW with the reference's paramwise weight decay,
    which also refreshes the bf16 operand copies (row f3, optim.py)

The detector heads (RPN / RoI / mask head, mmcv ops) are out of scope and mmcv/mmdet are not installed, so the
"loss" is sum_i <lateral_i, w_i> with fixed random w_i, as in bench.py.  Forward + backward + all-reduce replay from ONE CUDA
graph; the optimizer step runs outside it every step (its lr / bias corrections are host values).  Same timing rules as
bench.py: >= 3 warm-up steps, CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.

    python tools/train_step_bench.py [--steps K] [--warmup W] [--model swin_t|swin_b]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn as nn

IMG_HW = (800, 1333)
MODELS = {"swin_t": (dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24]), 0.1),
          "swin_b": (dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32]), 0.3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--model", default="swin_t", choices=sorted(MODELS))
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line (as bench.py): libraries that write to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    import swin_b200
    from swin_b200 import ops
    from swin_b200.ddp import BucketedGradAllReduce
    from swin_b200.fpn import SwinFPNLaterals
    from swin_b200.graph import GraphedStep
    from swin_b200.optim import FusedAdamW

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas, opts.config.min_ctas = 4, 1
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        from swin_b200 import _lib
        _lib.lib().swin_sm_reserve(4)
    cfg, dpr = MODELS[args.model]
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    torch.manual_seed(0)
    net = swin_b200.SwinTransformer(drop_path_rate=dpr, compute_dtype="bf16", **cfg)
    net.init_weights()
    widths = [cfg["embed_dim"] * 2 ** i for i in range(4)]
    model = SwinFPNLaterals(net, [nn.Conv2d(c, 256, 1) for c in widths]).to(dev).train()      # configs/_base_/models/mask_rcnn_swin_fpn.py:21-25
    ddp = BucketedGradAllReduce(model, bucket_mb=32.0)
    opt = FusedAdamW(model, lr=1e-4, betas=(0.9, 0.999), weight_decay=0.05)                   # configs/swin/..._1x_coco.py:64-67
    host = torch.from_numpy(np.random.default_rng(rank).standard_normal((B, 3) + IMG_HW).astype(np.float32)).pin_memory()
    x_dev = host.to(dev)
    with torch.no_grad():
        shapes = [tuple(o.shape) for o in model(x_dev[:1])]
    # cotangents in the laterals' own (channels-last) memory format
    cots_nhwc = [torch.randn((B, s[2], s[3], s[1]), device=dev) for s in shapes]
    cots = [c.permute(0, 3, 1, 2) for c in cots_nhwc]

    def fwd_bwd():
        ddp.zero_grad()
        outs = model(x_dev)
        with torch.no_grad():
            loss = sum(torch.dot(o.permute(0, 2, 3, 1).reshape(-1), c.reshape(-1)) for o, c in zip(outs, cots_nhwc))
        torch.autograd.backward(outs, cots)
        ddp.finish()
        return loss

    for _ in range(W):
        fwd_bwd()
        opt.step()
    l0 = ops.LAUNCHES
    fwd_bwd()
    opt.step()
    launches_per_step = ops.LAUNCHES - l0
    graphed = GraphedStep(fwd_bwd, warmup=1)

    def step():
        loss = graphed.replay()
        opt.step()
        return loss

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(2):
        step()
    ms = timed(step, K)
    ms_fb = timed(graphed.replay, K)                 # forward + backward (+ all-reduce) alone, same process, same box
    # the optimizer alone: host time to enqueue one step, and the GPU time of its launches (CUDA events around each one)
    import time
    from bench import KernelTimer
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    for _ in range(K):
        opt.step()
    host_opt = (time.perf_counter() - h0) * 1e3 / K
    torch.cuda.synchronize()
    timer = KernelTimer()
    ops.set_kernel_timer(timer)
    for _ in range(K):
        opt.step()
    ops.set_kernel_timer(None)
    n_opt, ms_opt, _, by_opt = timer.summary("adamw_step")
    loss = float(step().item())
    assert np.isfinite(loss), "non-finite loss"
    nparam = sum(p.numel() for p in model.parameters())
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps({
            "metric": f"{args.model}_backbone_fpn_laterals_adamw_train_step_images_per_sec_800x1333", "value": world * B * K / (ms / 1e3),
            "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "fwd_bwd_ms_per_step": ms_fb / K,
            "adamw": {"gpu_ms_per_step": ms_opt / K, "launches_per_step": n_opt // K, "host_enqueue_ms_per_step": host_opt,
                      "gb_per_s": by_opt / (ms_opt / 1e3) / 1e9 if ms_opt > 0 else None},
            "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic", "gpu_launches": launches_per_step * K,
            "config": {"workload": "configs[2] restricted to the rows in scope: backbone fwd+bwd + norm{i}/FPN lateral 1x1 convs (f2) + "
                                   "gradient all-reduce + fused AdamW step (f3); detector heads out of scope",
                       "per_gpu_batch": B, "global_batch": world * B, "parameters": nparam, "dispatch": "cuda_graph + optimizer outside the graph"}}))
    sys.stdout.flush()
    if world > 1:
        import gc
        import threading
        dist.barrier()
        torch.cuda.synchronize()
        graphed = None
        gc.collect()
        killer = threading.Timer(20.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        dist.destroy_process_group()
        killer.cancel()


if __name__ == "__main__":
    main()
