"""Micro-benchmark of the tcgen05 GEMM on the backbone's real shapes (CUDA events, L2 flushed between runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swin_b200 import ops, _lib as L

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
only = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5


def run(name, M, N, K, epi=L.EPI_STORE, a_trans=False, b_trans=False, out_f32=False, do_flush=True):
    if only and only != name:
        return
    A = torch.randn((K, M) if a_trans else (M, K), device=dev).bfloat16()
    B = (torch.randn((K, N) if b_trans else (N, K), device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    kw = {}
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if (out_f32 or epi in (L.EPI_RESIDUAL, L.EPI_ATOMIC_ADD)) else torch.bfloat16)
    if epi == L.EPI_GELU:
        kw["out2"] = torch.empty_like(out)
    if epi == L.EPI_RESIDUAL:
        kw["aux"] = torch.randn(M, N, device=dev)
    if epi == L.EPI_DGELU:
        kw["aux"] = torch.randn(M, N, device=dev).bfloat16()
    if epi == L.EPI_ATOMIC_ADD:
        out.zero_()
        bias = None
    ts = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(A, B, M, N, K, a_trans=a_trans, b_trans=b_trans, epilogue=epi, bias=bias, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2] * 1e-3
    byt = A.numel() * 2 + out.numel() * out.element_size() * (2 if epi == L.EPI_GELU else 1) + sum(v.numel() * v.element_size() for k, v in kw.items() if k == "aux")
    print(f"{name:28s} M={M:8d} N={N:5d} K={K:5d}  {t*1e6:8.1f} us  {2*M*N*K/t/1e12:7.1f} TF/s  {byt/t/1e9:7.0f} GB/s")


T0, Tp0 = 1068800, 1091328
T2, Tp2 = 67200, 75264
run("qkv_s0", Tp0, 288, 96)
run("proj_plain_s0", Tp0, 96, 96)
run("fc1_gelu_s0", T0, 384, 96, L.EPI_GELU)
run("fc1_store_s0", T0, 384, 96)
run("fc2_resid_s0", T0, 96, 384, L.EPI_RESIDUAL)
run("dgelu_s0", T0, 384, 96, L.EPI_DGELU, b_trans=True)
run("dxn_s0", T0, 96, 384, b_trans=True)
run("dW_fc1_s0", 384, 96, T0, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
run("qkv_s2", Tp2, 1152, 384)
run("fc1_gelu_s2", T2, 1536, 384, L.EPI_GELU)
run("fc1_store_s2", T2, 1536, 384)
run("fc2_resid_s2", T2, 384, 1536, L.EPI_RESIDUAL)
run("dgelu_s2", T2, 1536, 384, L.EPI_DGELU, b_trans=True)
run("dW_fc2_s2", 384, 1536, T2, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
run("dW_qkv_s2", 1152, 384, Tp2, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
run("dx_fc1_s2", T2, 384, 1536, b_trans=True)
run("dx_qkv_s2", Tp2, 384, 1152, b_trans=True)
run("dx_proj_s2", Tp2, 384, 384, b_trans=True)
run("square_8k", 8192, 8192, 8192)
run("square_4k_f32out", 4096, 4096, 4096, out_f32=True)
run("sq4k_tt", 4096, 4096, 4096, a_trans=True, b_trans=True)
run("dW_s2_small_flush", 384, 1536, 16800, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
run("dW_s2_small_L2hot", 384, 1536, 16800, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True, do_flush=False)
run("dW_fc1_s2", 1536, 384, T2, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
