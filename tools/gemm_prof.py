"""Where do the GEMM's warp roles wait?  Needs the development library (make -C .../csrc prof) via
SWIN_B200_LIB=.../libswin_b200_prof.so; run once with SWIN_GEMM_PAIR=0 and once with =1.

Per shape: kernel time and, for the first CTA (pair), the share of its lifetime each role spends blocked:
producer on slot-empty, MMA thread on accumulator-empty and on operand-full, epilogue warp 0 on accumulator-full."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swin_b200 import ops, _lib as L

dev = "cuda"
lib = L.lib()
buf = (C.c_ulonglong * 16)()
lib.swin_debug_gemm_prof.argtypes = [C.c_void_p, C.c_int]
lib.swin_debug_gemm_prof.restype = C.c_int
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def run(name, M, N, K, epi=L.EPI_STORE, a_trans=False, b_trans=False):
    A = torch.randn((K, M) if a_trans else (M, K), device=dev).bfloat16()
    B = (torch.randn((K, N) if b_trans else (N, K), device=dev) * 0.05).bfloat16()
    kw = {}
    f32 = epi in (L.EPI_RESIDUAL, L.EPI_ATOMIC_ADD)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    if epi == L.EPI_GELU:
        kw["out2"] = torch.empty_like(out)
    if epi == L.EPI_RESIDUAL:
        kw["aux"] = torch.randn(M, N, device=dev)
    if epi == L.EPI_DGELU:
        kw["aux"] = torch.randn(M, N, device=dev).bfloat16()
    for i in range(3):
        flush.zero_()
        lib.swin_debug_gemm_prof(C.byref(buf), 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if epi == L.EPI_ATOMIC_ADD and a_trans and b_trans:
            kw["colsum_a"] = torch.zeros(M, device=dev)
        ops.gemm(A, B, M, N, K, a_trans=a_trans, b_trans=b_trans, epilogue=epi, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
    lib.swin_debug_gemm_prof(C.byref(buf), 1)
    v = list(buf)
    us = e0.elapsed_time(e1) * 1e3
    pt, mt, et = max(v[2], 1), max(v[5], 1), max(v[15], 1)
    print(f"{name:14s} {us:8.1f} us {2*M*N*K/us/1e6:7.1f} TF/s | producer0 empty-wait {100*v[0]/pt:5.1f}%  producer1 {100*v[8]/max(v[10],1):5.1f}% | "
          f"MMA acc-empty {100*v[3]/mt:5.1f}% operand-full {100*v[4]/mt:5.1f}% issue {100*(mt-v[3]-v[4])/mt:5.1f}% | "
          f"epi0 acc-full rank0 {100*v[6]/max(mt,1):5.1f}% rank1 {100*v[14]/max(mt,1):5.1f}%  (MMA thread {mt/1e3:.0f} kclk; per k-block: "
          f"{mt/max(v[13],1):.0f} clk total, {v[11]/max(v[13],1):.0f} issuing MMAs, {v[12]/max(v[13],1):.0f} issuing commits, k-blocks {v[13]})", flush=True)


print("SWIN_GEMM_PAIR =", os.environ.get("SWIN_GEMM_PAIR", "(default)"))
T0, Tp0, T2, Tp2 = 1068800, 1091328, 67200, 75264
run("qkv_s0", Tp0, 288, 96)
run("proj_s0", Tp0, 96, 96)
run("fc2_resid_s0", T0, 96, 384, L.EPI_RESIDUAL)
run("qkv_s2", Tp2, 1152, 384)
run("fc1_store_s2", T2, 1536, 384)
run("fc1_gelu_s2", T2, 1536, 384, L.EPI_GELU)
run("fc2_resid_s2", T2, 384, 1536, L.EPI_RESIDUAL)
run("dgelu_s2", T2, 1536, 384, L.EPI_DGELU, b_trans=True)
run("dW_fc1_s2", 1536, 384, T2, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
run("dW_fc2_s2", 384, 1536, T2, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
run("dW_qkv_s2", 1152, 384, Tp2, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
run("dx_fc1_s2", T2, 384, 1536, b_trans=True)
run("dx_qkv_s2", Tp2, 384, 1152, b_trans=True)
run("square_8k", 8192, 8192, 8192)
