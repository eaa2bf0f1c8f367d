"""Staged correctness check of the tcgen05 GEMM (run with SWIN_GEMM_PAIR=2 to force the CTA-pair kernels wherever legal).

Cases go from one work unit to the backbone's real shapes; every case prints its relative error against an fp64
reference and, when it fails, the error broken down by 128-row block and by N-half (the two things the cta_group::2
layout splits), so a protocol or layout mistake is visible from one run.  Stops at the first CUDA error.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swin_b200 import ops, _lib as L

dev = "cuda"
bad = 0


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def breakdown(out, ref, N):
    o, r = out.double().cpu(), ref.double().cpu()
    M = o.shape[0]
    for m0 in range(0, min(M, 1024), 128):
        parts = []
        for h in range(4):
            c0, c1 = h * N // 4, (h + 1) * N // 4
            parts.append("%.1e" % rel(o[m0:m0 + 128, c0:c1], r[m0:m0 + 128, c0:c1]))
        print(f"      rows {m0:5d}+128  N-quarters: {' '.join(parts)}")


def case(name, M, N, K, epi=L.EPI_STORE, a_trans=False, b_trans=False, out_f32=True, tol=None):
    global bad
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn((K, M) if a_trans else (M, K), generator=g) * 0.5).bfloat16().to(dev)
    B = (torch.randn((K, N) if b_trans else (N, K), generator=g) * 0.5).bfloat16().to(dev)
    Ad = A.double().cpu().t() if a_trans else A.double().cpu()
    Bd = B.double().cpu() if b_trans else B.double().cpu().t()
    acc = Ad @ Bd
    bias = torch.randn(N, generator=g).to(dev)
    kw = {}
    want2 = None
    if epi == L.EPI_STORE:
        out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
        want = acc + bias.double().cpu()
    elif epi == L.EPI_GELU:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw["out2"] = torch.empty_like(out)
        u = (acc + bias.double().cpu()).requires_grad_(True)
        want = 0.5 * u * (1 + torch.erf(u / 2 ** 0.5))
        want.sum().backward()
        want, want2 = want.detach(), u.grad
    elif epi == L.EPI_DGELU:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw["aux"] = torch.randn(M, N, generator=g).bfloat16().to(dev)
        want = (acc + bias.double().cpu()) * kw["aux"].double().cpu()
    elif epi == L.EPI_RESIDUAL:
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
        kw["aux"] = torch.randn(M, N, generator=g).to(dev)
        want = kw["aux"].double().cpu() + acc + bias.double().cpu()
    else:  # ATOMIC_ADD (+ column sums of A^T = bias gradient)
        out = torch.zeros(M, N, device=dev, dtype=torch.float32)
        bias = None
        if a_trans and b_trans:
            kw["colsum_a"] = torch.zeros(M, device=dev)
        want = acc
    ops.gemm(A, B, M, N, K, a_trans=a_trans, b_trans=b_trans, epilogue=epi, bias=bias, out=out, **kw)
    torch.cuda.synchronize()
    e = rel(out, want)
    t = tol if tol is not None else (1e-5 if out.dtype == torch.float32 else 6e-3)
    msg = f"{name:34s} M={M:8d} N={N:5d} K={K:6d}  rel={e:.2e}"
    ok = e < t
    if want2 is not None:
        e2 = rel(kw["out2"], want2)
        msg += f" d={e2:.2e}"
        ok = ok and e2 < t
    if "colsum_a" in kw:
        e3 = rel(kw["colsum_a"], Ad.sum(1))
        msg += f" colsum={e3:.2e}"
        ok = ok and e3 < 1e-5
    print(msg, "ok" if ok else "FAIL", flush=True)
    if not ok:
        bad += 1
        breakdown(out, want, N)


print("SWIN_GEMM_PAIR =", os.environ.get("SWIN_GEMM_PAIR", "(default 1)"), flush=True)
case("one unit f32 store", 256, 256, 64)
case("one unit K=512", 256, 256, 512)
case("one unit N=128", 256, 128, 256)
case("one unit N=96 (half 48)", 256, 96, 192)
case("one unit N=16", 256, 16, 128)
case("tiles 4x2", 1024, 512, 512)
case("ragged M", 1000, 512, 320)
case("odd tile count", 1100, 192, 96)
case("bf16 store (class 1)", 1024, 512, 512, out_f32=False)
case("bf16 store N=192", 1024, 384, 384, out_f32=False)
case("B MN-major", 1024, 512, 512, b_trans=True)
case("B MN-major N=384 (bn 128)", 1024, 384, 512, b_trans=True, out_f32=False)
case("A MN-major", 1024, 256, 512, a_trans=True)
case("A,B MN-major", 1024, 256, 512, a_trans=True, b_trans=True)
case("gelu", 1024, 512, 128, L.EPI_GELU)
case("dgelu Bt", 1024, 512, 128, L.EPI_DGELU, b_trans=True)
case("residual", 1024, 128, 512, L.EPI_RESIDUAL)
case("residual N=384", 2048, 384, 1536, L.EPI_RESIDUAL)
case("split-K dW + colsum", 512, 256, 20000, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
case("split-K dW N=384", 1536, 384, 16800, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
case("split-K dW M=1152 (9 tiles)", 1152, 384, 18816, L.EPI_ATOMIC_ADD, a_trans=True, b_trans=True)
# backbone shapes (stage 2 / stage 0)
case("qkv s2", 75264, 1152, 384, out_f32=False)
case("fc1 gelu s3-rows", 16800, 1536, 384, L.EPI_GELU)
case("fc2 residual s2", 33600, 384, 1536, L.EPI_RESIDUAL)
case("dgelu s2", 16800, 1536, 384, L.EPI_DGELU, b_trans=True)
case("dx fc1 s2", 33600, 384, 1536, b_trans=True, out_f32=False)
case("qkv s0", 200000, 288, 96, out_f32=False)
case("fc2 residual s0", 200000, 96, 384, L.EPI_RESIDUAL)
print("FAILED CASES:", bad, flush=True)
sys.exit(1 if bad else 0)
