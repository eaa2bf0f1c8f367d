"""Micro-benchmark of the HBM-bound kernels (LayerNorm family, gathers, casts) on the backbone's real shapes
(B=16, 800x1333): CUDA events, L2 flushed between runs, algorithmic GB/s against the measured HBM peak."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swin_b200 import ops, _lib as L

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
only = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B = 16
STAGES = [(200, 334, 96), (100, 167, 192), (50, 84, 384), (25, 42, 768)]


def timeit(name, fn, nbytes):
    if only and not name.startswith(only):
        return
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2] * 1e-3
    print(f"{name:32s} {t*1e6:8.1f} us  {nbytes/t/1e9:7.0f} GB/s  ({100*nbytes/t/6527.1e9:5.1f}% of 6527)")


for si, (H, W, C) in enumerate(STAGES):
    T = B * H * W
    x = torch.randn(B, H * W, C, device=dev)
    g = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
    s = torch.rand(B, device=dev)
    for shift in (0, 3):
        y, mean, rstd = ops.ln_fwd(1, x, g, b, B, H, W, C, 7, shift, 1e-5, L.BF16)
        timeit(f"ln_fwd_gather_s{si}_shift{shift}", lambda: ops.ln_fwd(1, x, g, b, B, H, W, C, 7, shift, 1e-5, L.BF16), ops._nb(x, y))
        dyw = torch.randn(y.shape, device=dev).bfloat16()
        dres = torch.randn_like(x)
        timeit(f"ln_bwd_gather_s{si}_shift{shift}", lambda: ops.ln_bwd(1, dyw, x, g, mean, rstd, dres, B, H, W, C, 7, shift), ops._nb(dyw, x, dres, x))
    y0, mean0, rstd0 = ops.ln_fwd(0, x, g, b, B, H, W, C, 1, 0, 1e-5, L.BF16)
    timeit(f"ln_fwd_plain_s{si}", lambda: ops.ln_fwd(0, x, g, b, B, H, W, C, 1, 0, 1e-5, L.BF16), ops._nb(x, y0))
    dy0 = torch.randn(T, C, device=dev).bfloat16()
    dres = torch.randn_like(x)
    out = ops.ln_bwd(0, dy0, x, g, mean0, rstd0, dres, B, H, W, C, 1, 0, emit_windows=(7, 3, s))
    timeit(f"ln_bwd_plain_emit_s{si}", lambda: ops.ln_bwd(0, dy0, x, g, mean0, rstd0, dres, B, H, W, C, 1, 0, emit_windows=(7, 3, s)),
           ops._nb(dy0, x, dres, x, out[3]))
    timeit(f"scale_cast_s{si}", lambda: ops.scale_cast(x, s, 0, B, H, W, C, 1, 0, L.BF16, want_colsum=True), ops._nb(x) * 1.5)
    timeit(f"window_gather_s{si}", lambda: ops.window_gather(x, H, W, 7, 3), 2 * ops._nb(x))
    o, m2, r2 = ops.ln_nchw_fwd(x, g, b, H, W, 1e-5)
    timeit(f"ln_nchw_fwd_s{si}", lambda: ops.ln_nchw_fwd(x, g, b, H, W, 1e-5), ops._nb(x, o))
    do = torch.randn_like(o)
    timeit(f"ln_nchw_bwd_s{si}", lambda: ops.ln_nchw_bwd(do, x, g, m2, r2), ops._nb(do, x, x))
    if si < 3:
        g4 = torch.randn(4 * C, device=dev); b4 = torch.randn(4 * C, device=dev)
        ym, mm, rm = ops.ln_fwd(2, x, g4, b4, B, H, W, C, 1, 0, 1e-5, L.BF16)
        timeit(f"ln_fwd_merge_s{si}", lambda: ops.ln_fwd(2, x, g4, b4, B, H, W, C, 1, 0, 1e-5, L.BF16), ops._nb(x, ym))
        dym = torch.randn(ym.shape, device=dev).bfloat16()
        timeit(f"ln_bwd_merge_s{si}", lambda: ops.ln_bwd(2, dym, x, g4, mm, rm, None, B, H, W, C, 1, 0), ops._nb(dym, x, x))
img = torch.randn(B, 3, 800, 1333, device=dev)
cols = ops.patch_gather(img, 4, L.BF16)
timeit("patch_gather", lambda: ops.patch_gather(img, 4, L.BF16), ops._nb(img, cols))
