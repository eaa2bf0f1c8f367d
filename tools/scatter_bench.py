"""Micro-benchmark of the proj GEMM with the window-reverse scatter + drop-path + residual epilogue (REF:151, :237-252) on the
four Swin-T stage shapes (B=16, 800x1333): CUDA events, L2 flushed, algorithmic GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swin_b200 import ops, _lib as L

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
B = 16
for si, (H, W, C) in enumerate([(200, 334, 96), (100, 167, 192), (50, 84, 384), (25, 42, 768)]):
    if only is not None and si != only:
        continue
    Hp, Wp = -(-H // 7) * 7, -(-W // 7) * 7
    T, Tp = B * H * W, B * Hp * Wp
    o = torch.randn(Tp, C, device=dev).bfloat16()
    w = (torch.randn(C, C, device=dev) * 0.05).bfloat16()
    bias = torch.randn(C, device=dev)
    x = torch.randn(B, H * W, C, device=dev)
    s1 = torch.rand(B, device=dev)
    x1 = torch.empty_like(x)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(o, w, Tp, C, C, bias=bias, epilogue=L.EPI_SCATTER_RESIDUAL, out=x1, aux=x, row_scale=s1, geom=(H, W, 7, 3))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2] * 1e-3
    nb = o.numel() * 2 + 2 * x.numel() * 4
    print(f"proj_scatter_residual_s{si} M={Tp} N={C} K={C}  {t*1e6:8.1f} us  {2*Tp*C*C/t/1e12:6.1f} TF/s  {nb/t/1e9:6.0f} GB/s")
