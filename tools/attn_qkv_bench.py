"""Micro-benchmark: fused QKV + window attention kernel vs the qkv GEMM + stand-alone attention kernel (stage shapes of the
benchmark, B = 16 at 800x1333).  usage: attn_qkv_bench.py [reps] [ncases]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swin_b200 import ops

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
# stage 0 / 1 / 2 of the benchmark (B = 16, 800x1333: 29x48, 15x24, 8x12 windows per image) + Swin-B widths
cases = [(22272, 3, True), (22272, 3, False), (5760, 6, True), (1536, 12, True), (22272, 4, True), (5760, 8, True)]
GRID = {22272: (29, 48), 5760: (15, 24), 1536: (8, 12)}
if len(sys.argv) > 2:
    cases = cases[:int(sys.argv[2])]


def med(fn):
    ts = []
    for _ in range(reps):
        for _ in range(6): flush.zero_()      # L2 flush, long enough for the host to enqueue the timed launches behind it
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[reps // 2] * 1e-3


for B_, nH, masked in cases:
    C = nH * 32
    if not ops.window_attn_qkv_supported(C, nH, 7):
        print(f"C={C}: not supported"); continue
    xw = torch.randn(B_ * 49, C, device=dev).bfloat16()
    w = (torch.randn(3 * C, C, device=dev) / C ** 0.5).bfloat16()
    b = torch.randn(3 * C, device=dev) * 0.1
    bias = torch.randn(nH, 49, 49, device=dev) * 0.3
    mask = mnz = None
    canon = (0, 0)
    if masked:
        from oracle import swin_oracle as so
        canon = GRID[B_]
        mask = torch.from_numpy(so.shift_mask_np(canon[0] * 7, canon[1] * 7, 7, 3)).to(dev)
        mnz = ops.mask_nonzero(mask)
    sc = 32 ** -0.5
    t_gemm = med(lambda: ops.gemm(xw, w, B_ * 49, 3 * C, C, bias=b))
    qkv = ops.gemm(xw, w, B_ * 49, 3 * C, C, bias=b).view(B_, 49, 3 * C)
    t_attn = med(lambda: ops.window_attn_fwd(qkv, bias, mask, B_, nH, 7, sc, mnz, canon))
    t_fq = med(lambda: ops.window_attn_qkv_fwd(xw, w, b, bias, mask, B_, nH, 7, sc, mnz, canon, want_qkv=True))
    t_f = med(lambda: ops.window_attn_qkv_fwd(xw, w, b, bias, mask, B_, nH, 7, sc, mnz, canon, want_qkv=False))
    fl = 2.0 * B_ * 49 * C * 3 * C + 307328.0 * B_ * nH
    by = B_ * 49 * C * 2
    print(f"B_={B_} nH={nH} mask={masked!s:5s} | gemm {t_gemm*1e6:7.1f} + attn {t_attn*1e6:7.1f} = {(t_gemm+t_attn)*1e6:7.1f} us ({fl/(t_gemm+t_attn)/1e12:6.1f} TF/s) | "
          f"fused+qkv {t_fq*1e6:7.1f} us ({fl/t_fq/1e12:6.1f} TF/s, {5*by/t_fq/1e9:5.0f} GB/s) | fused {t_f*1e6:7.1f} us ({fl/t_f/1e12:6.1f} TF/s = {fl/t_f/1e12/1414.9*100:4.1f}% of 1414.9, {2*by/t_f/1e9:5.0f} GB/s)")
