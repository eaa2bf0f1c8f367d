"""ln_nchw_fwd / ln_nchw_bwd (output norm + NCHW store) on the Swin-T and Swin-B stage shapes (B=16, 800x1333): CUDA events,
L2 flushed between runs, algorithmic GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swin_b200 import ops

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
B = 16


def timeit(name, fn, nbytes):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2] * 1e-3
    print(f"{name:34s} {t*1e6:8.1f} us  {nbytes/t/1e9:7.0f} GB/s", flush=True)


for (H, W) in ((200, 334), (100, 167), (50, 84), (25, 42)):
    for C0 in (96, 128):
        C = C0 * (200 // H)
        x = torch.randn(B, H * W, C, device=dev)
        g = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
        o, m, r = ops.ln_nchw_fwd(x, g, b, H, W, 1e-5)
        timeit(f"ln_nchw_fwd C={C}", lambda: ops.ln_nchw_fwd(x, g, b, H, W, 1e-5), ops._nb(x, o))
        do = torch.randn_like(o)
        timeit(f"ln_nchw_bwd C={C}", lambda: ops.ln_nchw_bwd(do, x, g, m, r), ops._nb(do, x, x))
        del x, o, do
