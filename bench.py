#!/usr/bin/env python
"""Benchmark of record for the Swin backbone hot path (BASELINE.json):
Swin-T backbone fwd+bwd images/sec @800x1333 bf16, per-GPU batch 16, data-parallel over N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference algorithm on the host CPU cores (oracle port)

One "step" = one forward + backward of the backbone over one synthetic batch (loss = sum_i <out_i, w_i> with
fixed random cotangents), plus for N>1 the bucketed NCCL gradient all-reduce overlapped with backward.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SWIN_T = dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7)
SWIN_B = dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=7)
IMG_HW = (800, 1333)
# model -> (constructor kwargs, fwd+bwd GFLOP per image (SURVEY.md §8(d) / BASELINE.md §2), drop_path_rate of its config,
#           metric name, workload description)
MODELS = {
    "swin_t": (SWIN_T, 593.18, 0.1, "swin_t_backbone_fwd_bwd_images_per_sec_800x1333",
               "configs[1]: Swin-T backbone fwd+bwd bf16, batch 16 per GPU, synthetic 3x800x1333, window 7 with shift masks"),
    "swin_b": (SWIN_B, 2052.6, 0.3, "swin_b_backbone_fwd_bwd_images_per_sec_800x1333",
               "configs[3]: Cascade Mask R-CNN Swin-B backbone (embed 128, depths 2-2-18-2) fwd+bwd bf16, batch 16 per GPU, synthetic "
               "3x800x1333, NCCL gradient all-reduce (347 MB fp32) overlapped with backward"),
}
GFLOP_PER_IMG = MODELS["swin_t"][1]
METRIC = MODELS["swin_t"][3]
WORKLOAD = MODELS["swin_t"][4]


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return d, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class KernelTimer:
    """CUDA-event pairs around every library launch of an instrumented eager pass, on the launching (current) stream.
    Each record carries the kernel's algorithmic flops and bytes (ops.py), so every class gets its own roofline."""

    def __init__(self):
        self.recs = []          # (kind, flops, bytes, e0, e1)

    def begin(self, kind, flops=0.0, nbytes=0.0):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        self.recs.append((kind, flops, nbytes, e0, e1))

    def end(self):
        self.recs[-1][4].record()

    def summary(self, prefix="gemm_tc"):
        """(launches, ms, flops, bytes) summed over the records whose kind starts with ``prefix``."""
        torch.cuda.synchronize()
        sel = [r for r in self.recs if r[0].startswith(prefix)]
        return len(sel), sum(r[3].elapsed_time(r[4]) for r in sel), sum(r[1] for r in sel), sum(r[2] for r in sel)

    def own_roofline(self, prefix, peak_tflops, peak_gbs):
        """(sum over launches of max(flops/peak, bytes/peak) in ms, measured ms) for the kinds starting with ``prefix``: each
        launch is held against the bound that applies to IT (the K = 96 GEMMs of stage 0 are HBM kernels, the stage-2 ones
        tensor kernels), which one class-wide TFLOP/s figure cannot express."""
        torch.cuda.synchronize()
        sel = [r for r in self.recs if r[0].startswith(prefix)]
        roof = sum(max(r[1] / (peak_tflops * 1e12), r[2] / (peak_gbs * 1e9)) for r in sel) * 1e3
        return roof, sum(r[3].elapsed_time(r[4]) for r in sel)

    def table(self, steps, peak_tflops, peak_gbs):
        """Per kernel kind: launches/step, ms/step, roofline ms/step = max(flops/peak, bytes/peak), achieved rates."""
        torch.cuda.synchronize()
        agg = {}
        for kind, fl, nb, e0, e1 in self.recs:
            a = agg.setdefault(kind, [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += fl; a[3] += nb
        rows = []
        for kind, (n, ms, fl, nb) in agg.items():
            roof = max(fl / (peak_tflops * 1e12), nb / (peak_gbs * 1e9)) * 1e3
            rows.append((ms / steps, kind, n / steps, roof / steps, fl / (ms * 1e-3) / 1e12 if ms else 0.0, nb / (ms * 1e-3) / 1e9 if ms else 0.0))
        rows.sort(reverse=True)
        tot, troof = sum(r[0] for r in rows), sum(r[3] for r in rows)
        out = [f"{'ms/step':>8s} {'roof ms':>8s} {'gap ms':>7s} {'n':>4s} {'TF/s':>7s} {'GB/s':>6s}  kernel (eager pass, CUDA events; roof = max(flops/{peak_tflops:.0f} TF/s, bytes/{peak_gbs:.0f} GB/s))"]
        for ms, kind, n, roof, tf, gb in rows:
            out.append(f"{ms:8.3f} {roof:8.3f} {ms - roof:7.3f} {n:4.0f} {tf:7.1f} {gb:6.0f}  {kind}")
        out.append(f"{tot:8.3f} {troof:8.3f} {tot - troof:7.3f}       total of instrumented launches")
        return "\n".join(out)


def eager_gpu_step_fn(batch: int, dev, cfg):
    """The GPU bar: the reference algorithm as plain torch-eager ops ON THE SAME B200 under
    torch.autocast(bf16) -- the oracle port with torch's fused layer_norm / gelu primitives, i.e. the kernels the reference
    module itself would launch.  Baseline arm only: never on the product path, never in --impl reference."""
    from oracle import swin_oracle as so
    so.FAST_OPS = True
    shapes = so.param_shapes(**cfg)
    params = {k: v.to(dev).requires_grad_(True) for k, v in so.seeded_params(shapes, seed=0, noisy=False).items()}
    img = torch.from_numpy(np.random.default_rng(0).standard_normal((batch, 3) + IMG_HW).astype(np.float32)).to(dev)
    state = {"cots": None}

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = so.backbone_forward(img, params, **cfg)
        if state["cots"] is None:
            state["cots"] = [torch.randn(o.shape, device=dev, dtype=o.dtype) for o in outs]
        torch.autograd.backward(outs, state["cots"])
        for p in params.values():
            p.grad = None
    return step


def time_eager_gpu(dev, cfg, batch: int, steps: int, warmup: int):
    """(images/s, ms/step, batch actually used) of the eager-GPU arm; halves the batch on OOM."""
    b = batch
    while b >= 1:
        try:
            step = eager_gpu_step_fn(b, dev, cfg)
            for _ in range(max(1, warmup)):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            return b / (ms / 1e3), ms, b
        except torch.OutOfMemoryError:
            step = None
            torch.cuda.empty_cache()
            b //= 2
    return None, None, 0


def run_torch_eager(args, emit=print):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, gflop, _, metric, workload = MODELS[args.model]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    v, ms, b = time_eager_gpu(dev, cfg, args.batch, args.steps, max(args.warmup, 1))
    emit(json.dumps({
        "impl": "torch_eager", "metric": metric, "value": v, "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "per_gpu_batch": b,
                   "note": "oracle port of the reference algorithm as torch-eager CUDA ops under torch.autocast(bf16), on this GPU"}}))


def oracle_step_fn(batch: int, cfg=None):
    """The reference algorithm on the CPU (oracle port of mmdet/models/backbones/swin_transformer.py), fp32."""
    from oracle import swin_oracle as so
    SWIN_T = cfg or MODELS["swin_t"][0]
    shapes = so.param_shapes(**SWIN_T)
    params = {k: v.requires_grad_(True) for k, v in so.seeded_params(shapes, seed=0, noisy=False).items()}
    img = torch.from_numpy(np.random.default_rng(0).standard_normal((batch, 3) + IMG_HW).astype(np.float32))
    cots = None

    def step():
        nonlocal cots
        outs = so.backbone_forward(img, params, **SWIN_T)
        if cots is None:
            cots = [torch.from_numpy(np.random.default_rng(10 + i).standard_normal(tuple(o.shape)).astype(np.float32)) for i, o in enumerate(outs)]
        loss = sum((o * c).sum() for o, c in zip(outs, cots))
        loss.backward()
        for p in params.values():
            p.grad = None
        return float(loss)
    return step


def run_reference(args, emit=print):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, _, _, METRIC, WORKLOAD = MODELS[args.model]
    step = oracle_step_fn(1, cfg)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = 1.0 / dt
    sample = f"1 image 3x{IMG_HW[0]}x{IMG_HW[1]} fwd+bwd per step, fp32, {args.steps} steps"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "note": "CPU arm: bounded sample of the same workload"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))



def fused_attention_probe(dev, B, peak, hbm_peak):
    """The fused QKV-projection + window-attention forward kernel (swin_window_attn_qkv_fwd: the inference path of the attention
    branch; training keeps the GEMM + attention pair, which is as fast when q/k/v must be written for backward) on the window
    rows of stages 0-2 of this benchmark's geometry: CUDA events, median of 5, L2 flushed between runs.  achieved = algorithmic
    flops (2 T 3C C projection + 4 N^2 32 per (window, head) attention) / time."""
    import swin_b200  # noqa: F401
    from swin_b200 import ops
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {}
    for stage, (nH, gh, gw) in enumerate([(3, 29, 48), (6, 15, 24), (12, 8, 12)]):
        C, B_ = nH * 32, B * gh * gw
        if not ops.window_attn_qkv_supported(C, nH, 7):
            continue
        g = torch.Generator(device=dev).manual_seed(stage)
        xw = torch.randn(B_ * 49, C, device=dev, generator=g).bfloat16()
        w = (torch.randn(3 * C, C, device=dev, generator=g) / C ** 0.5).bfloat16()
        bq = torch.randn(3 * C, device=dev, generator=g) * 0.1
        bias = torch.randn(nH, 49, 49, device=dev, generator=g) * 0.3
        mask = ops.shift_mask(gh * 7, gw * 7, 7, 3, dev)
        mnz = ops.mask_nonzero(mask)
        ts = []
        for _ in range(5):
            for _ in range(4):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.window_attn_qkv_fwd(xw, w, bq, bias, mask, B_, nH, 7, 32 ** -0.5, mnz, (gh, gw), want_qkv=False, want_lse=False)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        t = sorted(ts)[2]
        fl = 2.0 * B_ * 49 * C * 3 * C + 307328.0 * B_ * nH
        by = 2.0 * B_ * 49 * C * 2
        out[f"stage{stage}_C{C}"] = {"windows": B_, "us": t * 1e6, "achieved": fl / t / 1e12, "unit": "TFLOP/s", "frac": fl / t / 1e12 / peak,
                                    "dram_gbs": by / t / 1e9, "dram_frac_of_hbm_peak": by / t / 1e9 / hbm_peak}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_eager"],
                    help="reference: the reference algorithm on the host CPU cores; torch_eager: the same algorithm as eager torch ops on the GPU")
    ap.add_argument("--model", default="swin_t", choices=sorted(MODELS), help="swin_t = BASELINE configs[1]; swin_b = configs[3]")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the torch-eager GPU baseline leg (N=1 only)")
    ap.add_argument("--attn-sweep", default=None, help="run the configs[4] window-attention sweep, write the table to this file, and exit")
    ap.add_argument("--sm-reserve", type=int, default=None, help="SMs the persistent kernels leave free for NCCL (default: 0 at N=1, "
                    "SWIN_SM_RESERVE or 4 at N>1)")
    ap.add_argument("--nccl-max-ctas", type=int, default=None, help="cap NCCL's CTAs per collective (default SWIN_NCCL_MAX_CTAS or 4)")
    ap.add_argument("--batch", type=int, default=16, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compute-dtype", default="bf16")
    ap.add_argument("--no-graph", action="store_true", help="eager dispatch instead of whole-step CUDA-graph replay")
    ap.add_argument("--breakdown", default=None, help="write the per-kernel roofline table of the instrumented pass to this file")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: warm up, then run exactly ONE step between cudaProfilerStart/Stop and exit "
                         "(use with ncu --profile-from-start off); prints no benchmark line")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 (NCCL prints its version banner there) are sent to
    # stderr for the duration of the run, and the real stdout is restored for the final print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str) -> None:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(line)
        sys.stdout.flush()
        os.dup2(2, 1)

    if args.impl == "reference":
        return run_reference(args, emit)
    if args.impl == "torch_eager":
        return run_torch_eager(args, emit)
    if args.attn_sweep:
        from tools.attn_sweep import run_sweep
        return run_sweep(args.attn_sweep)
    SWIN_T, GFLOP_PER_IMG, DPR, METRIC, WORKLOAD = MODELS[args.model]

    import torch.distributed as dist
    import swin_b200
    from swin_b200 import ops
    from swin_b200.ddp import BucketedGradAllReduce

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nccl_ctas = sm_reserve = 0
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The gradient all-reduce is latency-, not bandwidth-bound on NVSwitch (110 MB per step against >= 25 ms of compute):
        # cap NCCL at a few CTAs and leave exactly that many SMs out of the persistent kernels' grids, so a GEMM / attention
        # launch never waits a whole extra wave for the SMs NCCL is holding (the fixed +0.6..0.9 ms per step of round 1).
        nccl_ctas = args.nccl_max_ctas if args.nccl_max_ctas is not None else int(os.environ.get("SWIN_NCCL_MAX_CTAS", "4"))
        sm_reserve = args.sm_reserve if args.sm_reserve is not None else int(os.environ.get("SWIN_SM_RESERVE", str(nccl_ctas)))
        pg_opts = None
        if nccl_ctas > 0:
            try:
                pg_opts = dist.ProcessGroupNCCL.Options()
                pg_opts.config.max_ctas = nccl_ctas
                pg_opts.config.min_ctas = 1
            except Exception as e:
                sys.stderr.write(f"[bench] NCCL CTA cap unavailable ({e})\n")
                pg_opts, nccl_ctas = None, 0
        dist.init_process_group("nccl", device_id=dev, pg_options=pg_opts)
        if sm_reserve > 0:
            from swin_b200 import _lib
            _lib.lib().swin_sm_reserve(sm_reserve)
    W = max(args.warmup, 3)
    K = args.steps
    B = args.batch

    torch.manual_seed(0)
    net = swin_b200.SwinTransformer(drop_path_rate=DPR, compute_dtype=args.compute_dtype, **SWIN_T)
    net.init_weights()
    net = net.to(dev).train()
    ddp = BucketedGradAllReduce(net, bucket_mb=32.0)
    host = torch.from_numpy(np.random.default_rng(rank).standard_normal((B, 3) + IMG_HW).astype(np.float32)).pin_memory()
    x_dev = host.to(dev)
    with torch.no_grad():
        shapes = [tuple(o.shape) for o in net(x_dev[:1])]
    cots = [torch.randn((B,) + s[1:], device=dev) for s in shapes]

    def step(x):
        # loss = sum_i <out_i, w_i>: the scalar is computed with one fused dot per output (no product temporaries) and
        # backward is seeded with the cotangents w_i directly (d loss / d out_i == w_i)
        ddp.zero_grad()
        outs = net(x)
        with torch.no_grad():
            loss = sum(torch.dot(o.reshape(-1), c.reshape(-1)) for o, c in zip(outs, cots))
        torch.autograd.backward(outs, cots)
        ddp.finish()
        return loss

    host_ms = [0.0]

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        for _ in range(n):
            fn()
        host_ms[0] = (time.perf_counter() - h0) * 1e3 / max(n, 1)     # CPU time to ENQUEUE one step (no sync inside)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(W):
        step(x_dev)
    if args.ncu_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(x_dev)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # whole-step CUDA graph: one capture, then every timed step is a single graph launch (host cost ~0)
    l0 = ops.LAUNCHES
    step(x_dev)
    launches_per_step = ops.LAUNCHES - l0
    graphed = None
    if not args.no_graph:
        from swin_b200.graph import GraphedStep
        try:
            graphed = GraphedStep(lambda: step(x_dev), warmup=1)
        except Exception as e:                      # capture is an optimisation, never a requirement
            sys.stderr.write(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly\n")
            graphed = None
    run_step = (lambda: graphed.replay()) if graphed is not None else (lambda: step(x_dev))
    for _ in range(2):
        run_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(run_step, K)
    launches = launches_per_step * K
    ms_step = ms_total / K
    host_enqueue_ms = host_ms[0]
    value = world * B * K / (ms_total / 1e3)

    # end to end through the public API: every step uploads its batch from pinned host memory (H2D, 205 MB) and reads its
    # scalar loss back (D2H).  The upload of step i+1 runs on a copy stream while step i computes (the usual input
    # prefetch): it lands in a staging buffer and is moved into the graph's static input by a device-to-device copy.
    copy_stream = torch.cuda.Stream()
    x_stage = torch.empty_like(x_dev)
    ev_ready, ev_free = torch.cuda.Event(), torch.cuda.Event()

    def upload():
        copy_stream.wait_event(ev_free)                      # the staging buffer has been consumed
        with torch.cuda.stream(copy_stream):
            x_stage.copy_(host, non_blocking=True)
            ev_ready.record(copy_stream)

    # the loss of every step is read back into pinned host memory by an asynchronous D2H copy (4 bytes per step, inside the
    # timed region); the host looks at the values after the loop, as a training loop that logs its loss does -- a blocking
    # .item() per step would only add a host round trip between two graph replays
    loss_host = torch.zeros((K + 2,), dtype=torch.float32).pin_memory()
    e2e_i = [0]

    def e2e_step():
        cur = torch.cuda.current_stream()
        cur.wait_event(ev_ready)                             # this step's batch has arrived
        x_dev.copy_(x_stage, non_blocking=True)
        ev_free.record(cur)
        upload()                                             # next step's batch streams in while this one computes
        loss = run_step()
        loss_host[e2e_i[0] % (K + 2)].copy_(loss.detach().reshape(()), non_blocking=True)
        e2e_i[0] += 1
    ev_free.record(torch.cuda.current_stream())
    upload()
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    e2e_i[0] = 0
    ms_e2e = timed(e2e_step, K)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(loss_host[:K]).all()), "non-finite loss in the end-to-end loop"
    e2e_value = world * B * K / (ms_e2e / 1e3)

    # dominant kernel class (tcgen05 GEMMs: 92.7 % of the FLOPs) timed launch by launch in an instrumented pass
    timer = KernelTimer()
    ops.set_kernel_timer(timer)
    timed(lambda: step(x_dev), K)
    ops.set_kernel_timer(None)
    n_launch, gemm_ms, gemm_flops, _ = timer.summary("gemm_tc")
    clocks = sampler.stop() if rank == 0 else None

    peaks, peak_kind = measured_peaks()
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    if args.breakdown and rank == 0:
        with open(args.breakdown, "w") as f:
            f.write(timer.table(K, peak, float(peaks.get("hbm_gbs", 6650.0))) + "\n")
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    # DRAM bytes of the GEMM class from the committed ncu launch list of this same command (tools/profile_step.sh), per launch
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "gemm_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("launches_per_step") and B == 16 and args.compute_dtype == "bf16":
            traffic = tj["dram_bytes_per_step"] / tj["launches_per_step"]
            traffic_src = "profiles/gemm_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum over the 155 GEMM launches of one step)"
    except Exception:
        pass
    _, _, _, gemm_bytes = timer.summary("gemm_tc")
    # the window-attention kernels (the metric's "attn % peak"): every launch whose kind starts with "attn" -- the attention
    # core (attn_fwd / attn_bwd) and, where the fused QKV + attention kernel runs, that kernel with its projection flops
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    attn = {}
    for pref in ("attn_fwd", "attn_qkv_fwd", "attn_bwd", "attn"):
        n_a, ms_a, fl_a, by_a = timer.summary(pref)
        if n_a and ms_a > 0:
            tf = fl_a / (ms_a / 1e3) / 1e12
            attn[pref if pref != "attn" else "all"] = {
                "launches_per_step": n_a // max(K, 1), "ms_per_step": ms_a / max(K, 1), "achieved": tf, "unit": "TFLOP/s",
                "frac": tf / peak, "dram_gbs": by_a / (ms_a / 1e3) / 1e9, "dram_frac_of_hbm_peak": by_a / (ms_a / 1e3) / 1e9 / hbm_peak}
    if rank == 0 and world == 1 and args.compute_dtype == "bf16":
        try:
            attn["fused_qkv_fwd_probe"] = fused_attention_probe(dev, B, peak, hbm_peak)
        except Exception as e:                      # the probe must never take the benchmark line down
            attn["fused_qkv_fwd_probe"] = {"error": str(e)[:200]}
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 bf16 GEMM, all epilogues)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": gemm_bytes / max(n_launch, 1),
                # the class is HBM-bound in stages 0-1 and tensor-bound in stages 2-3: its DRAM rate beside the tensor rate
                "dram_gbs": (traffic * n_launch / (gemm_ms / 1e3) / 1e9) if (traffic and gemm_ms > 0) else None,
                "dram_frac_of_hbm_peak": (traffic * n_launch / (gemm_ms / 1e3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0))) if (traffic and gemm_ms > 0) else None,
                "peak_source": peak_kind + " (sustained: timed inside a long step)",
                "launches_per_step": n_launch // max(K, 1), "ms_per_step_in_kernel": gemm_ms / max(K, 1),
                "whole_step_frac_of_tensor_roofline": (GFLOP_PER_IMG * 1e9 * B / (ms_step / 1e3)) / (peak * 1e12),
                "attention": attn}
    # every launch against its own bound (max of the tensor time and the HBM time of its algorithmic flops / bytes)
    for key, pref in (("own_roofline_frac", "gemm_tc"), ("own_roofline_frac_all_kernels", "")):
        r_ms, t_ms = timer.own_roofline(pref, peak, hbm_peak)
        roofline[key] = (r_ms / t_ms) if t_ms > 0 else None

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cstep = oracle_step_fn(1, SWIN_T)
        t0 = time.perf_counter(); cstep(); warm = time.perf_counter() - t0
        n = 3 if warm < 8 else 1
        t0 = time.perf_counter()
        for _ in range(n):
            cstep()
        dt = (time.perf_counter() - t0) / n
        cpu_baseline = {"value": 1.0 / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"oracle port, fp32, 1 image 3x{IMG_HW[0]}x{IMG_HW[1]} fwd+bwd, 1 warm-up + {n} timed"}

    dispatch = "cuda_graph" if graphed is not None else "eager"
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager:
        # the GPU bar: same algorithm as torch-eager ops on this same GPU (bounded: 1 warm-up + 2 timed steps)
        graphed = None
        torch.cuda.empty_cache()
        try:
            v_e, ms_e, b_e = time_eager_gpu(dev, SWIN_T, B, 2, 1)
            if v_e:
                gpu_eager = {"value": v_e, "unit": "images/s", "ms_per_step": ms_e, "per_gpu_batch": b_e, "kind": "port",
                             "sample": "oracle port as torch-eager CUDA ops under torch.autocast(bf16) (fused layer_norm/gelu/softmax "
                                       "primitives), same GPU, 1 warm-up + 2 timed steps", "speedup_ours": value / v_e}
        except Exception as e:
            gpu_eager = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        from oracle import swin_oracle as _so
        _so.FAST_OPS = False

    if rank == 0:
        img_bytes = B * 3 * IMG_HW[0] * IMG_HW[1] * 4
        emit(json.dumps({
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.compute_dtype == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "per_gpu_batch": B, "drop_path_rate": DPR,
                       "nccl_max_ctas": nccl_ctas, "sm_reserve": sm_reserve,
                       "parallelism": f"dp{world}", "dispatch": dispatch, "l2": "inputs (205 MB/step) and activations (>10 GB/step) exceed the 126 MB L2",
                       "precision": "bf16 tcgen05 operands, fp32 accumulate/LN/softmax/residual stream, fp32 master weights"},
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": img_bytes, "d2h_bytes_per_step": 4,
                    "input_pipeline": "pinned host -> staging buffer on a copy stream (overlaps the previous step) -> device-to-device into the static input; "
                                      "loss: async D2H into pinned memory every step, checked after the loop"},
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "gpu_eager_baseline": gpu_eager}))
    sys.stdout.flush()
    if world > 1:
        # Clean teardown: the captured graph holds NCCL kernels of this communicator, so it is destroyed first (graph, its
        # private pool, the events), the device is drained, and only then is the process group torn down.  A watchdog turns a
        # teardown that still hangs (seen on torch 2.11 / NCCL 2.28 when the graph outlives the communicator) into exit 0.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        graphed = run_step = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        killer = threading.Timer(20.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        dist.destroy_process_group()
        killer.cancel()


if __name__ == "__main__":
    main()
