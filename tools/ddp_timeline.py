"""Where does the multi-GPU fixed cost go?  One data-parallel step of the benchmark (eager dispatch, so every kernel is a
separate launch) under torch.profiler (CUPTI) on every rank; rank 0 writes a compact timeline: per stream the busy time,
every NCCL kernel with its start relative to the step and how much of it overlaps compute, the exposed tail (all-reduce
time after the last compute kernel) and the gaps on the compute stream.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29700 \
        tools/ddp_timeline.py OUT.txt [--model swin_t] [--nccl-max-ctas 4] [--sm-reserve 4] [--graph]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--model", default="swin_t")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--nccl-max-ctas", type=int, default=4)
    ap.add_argument("--sm-reserve", type=int, default=4)
    ap.add_argument("--graph", action="store_true", help="profile CUDA-graph replays (as the benchmark runs) instead of eager steps")
    args = ap.parse_args()
    import bench
    import swin_b200
    from swin_b200 import _lib
    from swin_b200.ddp import BucketedGradAllReduce
    cfg, _, dpr, _, _ = bench.MODELS[args.model]
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        opts = None
        if args.nccl_max_ctas > 0:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = args.nccl_max_ctas
            opts.config.min_ctas = 1
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        if args.sm_reserve > 0:
            _lib.lib().swin_sm_reserve(args.sm_reserve)
    torch.manual_seed(0)
    net = swin_b200.SwinTransformer(drop_path_rate=dpr, **cfg)
    net.init_weights()
    net = net.to(dev).train()
    ddp = BucketedGradAllReduce(net, bucket_mb=32.0)
    x = torch.from_numpy(np.random.default_rng(rank).standard_normal((args.batch, 3) + bench.IMG_HW).astype(np.float32)).to(dev)
    with torch.no_grad():
        shapes = [tuple(o.shape) for o in net(x[:1])]
    cots = [torch.randn((args.batch,) + s[1:], device=dev) for s in shapes]

    def step():
        ddp.zero_grad()
        outs = net(x)
        torch.autograd.backward(outs, cots)
        ddp.finish()

    for _ in range(3):
        step()
    run = step
    if args.graph:
        from swin_b200.graph import GraphedStep
        g = GraphedStep(step, warmup=1)
        run = g.replay
        for _ in range(2):
            run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            run()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "stream", -1) if hasattr(e, "stream") else -1) for e in evs), key=lambda t: t[0])
    if rank == 0 and ks:
        is_nccl = lambda n: "nccl" in n.lower()
        comp = [k for k in ks if not is_nccl(k[2])]
        nccl = [k for k in ks if is_nccl(k[2])]
        # split the two profiled steps at the largest gap between compute kernels near the middle
        mid = len(comp) // 2
        t0, t_mid = comp[0][0], comp[mid][0]
        out = [f"# ddp timeline: {args.model}, world {world}, per-GPU batch {args.batch}, dispatch {'cuda_graph' if args.graph else 'eager'}, "
               f"nccl max_ctas {args.nccl_max_ctas}, sm_reserve {args.sm_reserve}; rank 0; two steps profiled, times in us from the first kernel",
               f"# compute kernels {len(comp)}, nccl kernels {len(nccl)}"]
        busy = sum(e - s for s, e, _, _ in comp)
        span = comp[-1][1] - comp[0][0]
        out.append(f"compute stream: span {span:.0f} us for 2 steps, kernel-busy {busy:.0f} us, idle inside the span {span - busy:.0f} us")
        # union of compute intervals for overlap accounting
        iv = sorted((s, e) for s, e, _, _ in comp)
        def overlap(a, b):
            tot = 0.0
            for s, e in iv:
                if e <= a: continue
                if s >= b: break
                tot += min(e, b) - max(s, a)
            return tot
        out.append(f"{'start':>10s} {'dur':>8s} {'overlapped':>10s}  nccl kernel")
        for s, e, n, _ in nccl:
            out.append(f"{s - t0:10.0f} {e - s:8.0f} {overlap(s, e):10.0f}  {n[:90]}")
        for label, lo, hi in (("step 1", t0, t_mid), ("step 2", t_mid, comp[-1][1] + 1e9)):
            cs = [k for k in comp if lo <= k[0] < hi]
            ns = [k for k in nccl if lo <= k[0] < hi]
            if cs and ns:
                tail = max(0.0, max(e for _, e, _, _ in ns) - max(e for _, e, _, _ in cs))
                out.append(f"{label}: compute {cs[0][0] - t0:.0f} .. {cs[-1][1] - t0:.0f} us, last nccl kernel ends {max(e for _, e, _, _ in ns) - t0:.0f} us -> exposed tail {tail:.0f} us; "
                           f"nccl busy {sum(e - s for s, e, _, _ in ns):.0f} us, of which under compute {sum(overlap(s, e) for s, e, _, _ in ns):.0f} us")
        gaps = sorted(((comp[i + 1][0] - max(c[1] for c in comp[max(0, i - 3):i + 1]), comp[i][2][:60], comp[i + 1][2][:60]) for i in range(len(comp) - 1)), reverse=True)[:12]
        out.append("largest gaps on the compute stream (us, after kernel -> before kernel):")
        for gp, a, b in gaps:
            out.append(f"  {gp:8.1f}  {a}  ->  {b}")
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            f.write("\n".join(out) + "\n")
        print("\n".join(out[:60]))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
