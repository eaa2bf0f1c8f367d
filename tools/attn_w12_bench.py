import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swin_b200 import ops
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def med(fn, reps=5):
    ts = []
    for _ in range(reps):
        for _ in range(4): flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[reps // 2] * 1e-3
for B_, nH in [(1000, 3), (10000, 3), (10000, 12), (2000, 24)]:
    N, C = 144, nH * 32
    qkv = torch.randn(B_, N, 3 * C, device=dev).bfloat16()
    bias = torch.randn(nH, N, N, device=dev) * 0.3
    o, lse = ops.window_attn_fwd(qkv, bias, None, B_, nH, 12, 32 ** -0.5)
    do = torch.randn_like(o)
    dbias = torch.zeros(nH, N, N, device=dev)
    tf = med(lambda: ops.window_attn_fwd(qkv, bias, None, B_, nH, 12, 32 ** -0.5))
    tb = med(lambda: ops.window_attn_bwd(qkv, o, do, lse, bias, None, B_, nH, 12, 32 ** -0.5, dbias=dbias))
    by = B_ * N * C * 2
    print(f"ws=12 B_={B_} nH={nH}: fwd {tf*1e6:8.1f} us {4*by/tf/1e9:6.0f} GB/s {4.0*B_*nH*N*N*32/tf/1e12:6.1f} TF/s | bwd {tb*1e6:8.1f} us {8*by/tb/1e9:6.0f} GB/s")
