"""Micro-benchmark of the fused window-attention kernels (BASELINE config 5 style sweep)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swin_b200 import ops

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cases = [(22272, 3, True), (22272, 3, False), (5760, 6, True), (1536, 12, True), (384, 24, True)]
if len(sys.argv) > 2:
    cases = cases[:int(sys.argv[2])]
for B_, nH, masked in cases:
    C = nH * 32
    qkv = torch.randn(B_, 49, 3 * C, device=dev).bfloat16()
    bias = torch.randn(nH, 49, 49, device=dev) * 0.3
    mask = mask_nz = None
    canon = (0, 0)
    if masked:                                   # the canonical SW-MSA mask of the benchmark's stage geometry (B = 16, 800x1333)
        canon = {22272: (29, 48), 5760: (15, 24), 1536: (8, 12), 384: (4, 6)}[B_]
        mask = ops.shift_mask(canon[0] * 7, canon[1] * 7, 7, 3, dev)
        mask_nz = ops.mask_nonzero(mask)
    dout = torch.randn(B_, 49, C, device=dev).bfloat16()
    tf, tb = [], []
    for _ in range(reps):
        for _ in range(6): flush.zero_()      # L2 flush, long enough for the host to enqueue the timed launches behind it
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        o, lse = ops.window_attn_fwd(qkv, bias, mask, B_, nH, 7, 32 ** -0.5, mask_nz, canon)
        e[1].record()
        dqkv, dbias = ops.window_attn_bwd(qkv, o, dout, lse, bias, mask, B_, nH, 7, 32 ** -0.5, mask_nz, canon)
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
    f, b = sorted(tf)[reps // 2] * 1e-3, sorted(tb)[reps // 2] * 1e-3
    wh = B_ * nH
    print(f"B_={B_:6d} nH={nH:2d} mask={masked!s:5s} fwd {f*1e6:8.1f} us {wh*12544/f/1e9:6.0f} GB/s {wh*307328/f/1e12:6.1f} TF/s | "
          f"bwd {b*1e6:8.1f} us {wh*21952/b/1e9:6.0f} GB/s {wh*768320/b/1e12:6.1f} TF/s")
