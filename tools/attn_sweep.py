"""BASELINE.json configs[4]: window-attention micro-benchmark sweep -- window 7 / 12, head_dim 32, shift mask on / off,
10^3 .. 10^5 windows -- our kernels against the reference's WindowAttention arithmetic as torch-eager CUDA ops under
torch.autocast(bf16) on the same GPU (the oracle port of mmdet/models/backbones/swin_transformer.py:121-153).

    python bench.py --attn-sweep profiles/r02/attn_sweep.txt        (or: python tools/attn_sweep.py OUT)

Columns: `core` = the attention kernel alone (qkv -> out), `module` = the whole WindowAttention.forward (qkv Linear +
attention + proj Linear), forward and forward+backward; TF/s counts the UNPADDED algorithmic flops."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def _time(fn, reps, flush):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e-3


def run_sweep(out_path: str, reps: int = 5):
    import swin_b200
    from swin_b200 import ops
    from oracle import swin_oracle as so
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    lines = [f"# window-attention sweep (configs[4]); GPU {torch.cuda.get_device_name(dev)}; median of {reps}, L2 flushed between runs",
             "# ours: bf16 tcgen05 kernels; eager: the reference arithmetic as torch CUDA ops under autocast(bf16); us = microseconds",
             f"{'ws':>3s} {'nH':>3s} {'windows':>7s} {'mask':>5s} | {'core fwd us':>11s} {'GB/s':>6s} {'TF/s':>6s} | {'core bwd us':>11s} {'GB/s':>6s} | "
             f"{'module fwd':>10s} {'eager fwd':>10s} {'x':>5s} | {'module f+b':>10s} {'eager f+b':>10s} {'x':>5s}"]
    for ws in (7, 12):
        N = ws * ws
        for nH in (3, 6, 12, 24):
            C = nH * 32
            for nwin in (1000, 10000, 100000):
                for masked in (False, True):
                    nW = 100 if masked else 0                     # B_ must be a multiple of nW
                    B_ = nwin
                    row = f"{ws:3d} {nH:3d} {B_:7d} {str(masked):>5s} | "
                    try:
                        torch.manual_seed(0)
                        m = swin_b200.WindowAttention(C, (ws, ws), nH, compute_dtype="bf16").to(dev)
                        with torch.no_grad():
                            m.relative_position_bias_table.normal_(0, 0.02)
                        mask = None
                        if masked:
                            side = 10
                            mask = torch.from_numpy(so.shift_mask_np(side * ws, side * ws, ws, ws // 2)).to(dev)
                        # ---- attention core
                        core_ok = True
                        try:
                            qkv = torch.randn(B_, N, 3 * C, device=dev).bfloat16()
                            bias = ops.rel_bias_expand(m.relative_position_bias_table.detach().contiguous(), ws)
                            mnz = ops.mask_nonzero(mask) if masked else None
                            canon = (10, 10) if (masked and ws == 7) else (0, 0)
                            o, lse = ops.window_attn_fwd(qkv, bias, mask, B_, nH, ws, 32 ** -0.5, mnz, canon)
                            dout = torch.randn_like(o)
                            tf = _time(lambda: ops.window_attn_fwd(qkv, bias, mask, B_, nH, ws, 32 ** -0.5, mnz, canon), reps, flush)
                            tb = _time(lambda: ops.window_attn_bwd(qkv, o, dout, lse, bias, mask, B_, nH, ws, 32 ** -0.5, mnz, canon), reps, flush)
                            fl = 4.0 * B_ * nH * N * N * 32
                            row += f"{tf * 1e6:11.1f} {B_ * N * C * 8 / tf / 1e9:6.0f} {fl / tf / 1e12:6.1f} | {tb * 1e6:11.1f} {B_ * N * C * 14 / tb / 1e9:6.0f} | "
                            del qkv, o, lse, dout
                        except RuntimeError as e:
                            core_ok = False
                            row += f"{'n/a':>11s} {'':>6s} {'':>6s} | {'n/a':>11s} {'':>6s} | "
                        # ---- whole module vs eager
                        x = torch.randn(B_, N, C, device=dev, requires_grad=True)
                        cot = torch.randn(B_, N, C, device=dev)
                        p = {"qkv.weight": m.qkv.weight.detach().clone().requires_grad_(True), "qkv.bias": m.qkv.bias.detach().clone().requires_grad_(True),
                             "proj.weight": m.proj.weight.detach().clone().requires_grad_(True), "proj.bias": m.proj.bias.detach().clone().requires_grad_(True),
                             "relative_position_bias_table": m.relative_position_bias_table.detach().clone().requires_grad_(True)}

                        def ours_f():
                            with torch.no_grad():
                                m(x, mask)

                        def ours_fb():
                            y = m(x, mask)
                            y.backward(cot)
                            x.grad = None
                            m.zero_grad(set_to_none=True)

                        def eager_f():
                            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                                so.window_attention(x, p, "", nH, ws, mask)

                        def eager_fb():
                            with torch.autocast("cuda", dtype=torch.bfloat16):
                                y = so.window_attention(x, p, "", nH, ws, mask)
                            y.backward(cot.to(y.dtype))
                            x.grad = None
                            for v in p.values():
                                v.grad = None
                        try:
                            ours_f(); ours_fb()
                            to_f, to_fb = _time(ours_f, reps, flush), _time(ours_fb, reps, flush)
                        except RuntimeError as e:
                            to_f = to_fb = None
                        try:
                            eager_f(); eager_fb()
                            te_f, te_fb = _time(eager_f, reps, flush), _time(eager_fb, reps, flush)
                        except torch.OutOfMemoryError:
                            te_f = te_fb = None
                            torch.cuda.empty_cache()
                        f = lambda t: f"{t * 1e6:10.1f}" if t else f"{'n/a':>10s}"
                        r = lambda a, b: f"{b / a:5.1f}" if (a and b) else f"{'':>5s}"
                        row += f"{f(to_f)} {f(te_f)} {r(to_f, te_f)} | {f(to_fb)} {f(te_fb)} {r(to_fb, te_fb)}"
                        del x, cot, p, m
                    except torch.OutOfMemoryError:
                        row += "out of memory"
                    torch.cuda.empty_cache()
                    lines.append(row)
                    print(row, file=sys.stderr, flush=True)
    text = "\n".join(lines) + "\n"
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    with open(out_path, "w") as fh:
        fh.write(text)
    return text


if __name__ == "__main__":
    run_sweep(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/attn_sweep.txt")
