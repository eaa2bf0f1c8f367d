"""Kernel-level parity on a B200: every C-ABI entry point against the oracle / plain torch math.
Index ops are bit-exact; fp32 kernels <=1e-5; bf16 tensor-core kernels are compared with the same
math done in fp64 on bf16-rounded inputs (tolerance 1e-2, the output rounding)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import swin_oracle as so
from oracle.make_golden import INDEX_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from swin_b200 import ops, _lib
    return ops, _lib


def rel(a, b):
    return so.rel_l2(a, b)


# ---------------------------------------------------------------- index ops
@pytest.mark.parametrize("ci", range(len(INDEX_CASES)))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_index_ops_bit_exact(ci, dtype):
    ops, _ = _ops()
    d = np.load(os.path.join(GOLDEN, "index_ops.npz"))
    B, H, W, C, ws = INDEX_CASES[ci]
    Hp, Wp = so.padded_hw(H, W, ws)
    # bf16 represents integers exactly only up to 256: use the value mod 251 for the bf16 run
    f = (lambda a: a) if dtype == torch.float32 else (lambda a: np.mod(a, 251))
    x = np.random.default_rng(100 + ci).integers(-30000, 30000, size=(B, Hp, Wp, C)).astype(np.float32)
    xt = torch.from_numpy(f(x)).to(DEV, dtype)
    part = ops.window_partition(xt, ws)
    assert torch.equal(part.float().cpu(), torch.from_numpy(f(d[f"part{ci}"].astype(np.float32))))
    assert torch.equal(ops.window_reverse(part, ws, Hp, Wp), xt)
    xs = np.random.default_rng(200 + ci).integers(1, 30000, size=(B, H, W, C)).astype(np.float32)
    for shift in (0, ws // 2):
        g = d[f"gather{ci}_s{shift}"].astype(np.float32)
        got = ops.window_gather(torch.from_numpy(f(xs)).to(DEV, dtype).reshape(B, H * W, C), H, W, ws, shift)
        want = torch.from_numpy(np.where(g == 0, 0, f(g)))          # pad slots stay exactly 0
        assert torch.equal(got.float().cpu(), want)
        ww = np.random.default_rng(300 + ci + shift).integers(1, 30000, size=g.shape).astype(np.float32)
        back = ops.window_scatter(torch.from_numpy(f(ww)).to(DEV, dtype), B, H, W, ws, shift)
        assert torch.equal(back.float().cpu(), torch.from_numpy(f(d[f"scatter{ci}_s{shift}"].astype(np.float32))))
    m = ops.shift_mask(H, W, ws, ws // 2, DEV)
    assert torch.equal(m.cpu(), torch.from_numpy(d[f"mask{ci}"]))


def test_gather_scatter_roundtrip_full_size():
    """cfg2 stage-0 geometry (B=2 of 16): scatter(gather(x)) == x bit-for-bit, pad slots are zero."""
    ops, _ = _ops()
    B, H, W, C, ws = 2, 200, 334, 96, 7
    x = torch.randn(B, H * W, C, device=DEV).bfloat16()
    for shift in (0, 3):
        xw = ops.window_gather(x, H, W, ws, shift)
        assert torch.equal(ops.window_scatter(xw, B, H, W, ws, shift), x)
        idx = torch.from_numpy(so.gather_index(H, W, ws, shift)).to(DEV)
        flat = xw.reshape(B, -1, C)
        assert flat[:, idx < 0].abs().max().item() == 0
        assert torch.equal(flat[:, idx >= 0], x[:, idx[idx >= 0]])


def test_rel_bias_expand_reduce():
    ops, _ = _ops()
    for ws, nH in ((7, 3), (12, 4)):
        table = torch.randn((2 * ws - 1) ** 2, nH, device=DEV)
        idx = torch.from_numpy(so.relative_position_index_np(ws)).to(DEV)
        want = table[idx.reshape(-1)].reshape(ws * ws, ws * ws, nH).permute(2, 0, 1).contiguous()
        assert torch.equal(ops.rel_bias_expand(table, ws), want)
        db = torch.randn(nH, ws * ws, ws * ws, device=DEV)
        ref = torch.zeros_like(table).index_put_((idx.reshape(-1),), db.permute(1, 2, 0).reshape(-1, nH), accumulate=True)
        assert rel(ops.rel_bias_reduce(db, ws), ref) < 1e-5


# ---------------------------------------------------------------- LayerNorm family
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("ydt", ["f32", "bf16"])
@pytest.mark.parametrize("C", [32, 96, 384])
def test_layernorm_modes(mode, ydt, C):
    ops, L = _ops()
    B, H, W, ws, shift = 2, 9, 13, 7, 3
    code = L.F32 if ydt == "f32" else L.BF16
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((B, H * W, C)).astype(np.float32))
    width = 4 * C if mode == 2 else C
    gm = torch.from_numpy((1 + 0.2 * rng.standard_normal(width)).astype(np.float32))
    bt = torch.from_numpy((0.3 * rng.standard_normal(width)).astype(np.float32))
    xr = x.double().requires_grad_(True)
    gr, br = gm.double().requires_grad_(True), bt.double().requires_grad_(True)
    if mode == 0:
        yr = so.layer_norm(xr, gr, br)
    elif mode == 1:
        yr = so.shift_gather(so.layer_norm(xr, gr, br), H, W, ws, shift)
    else:
        idx = torch.from_numpy(so.merge_index(H, W))
        xz = torch.cat([xr, xr.new_zeros(B, 1, C)], 1)
        sel = torch.where(idx < 0, torch.full_like(idx, H * W), idx)
        yr = so.layer_norm(xz[:, sel.reshape(-1)].reshape(B, idx.shape[0], 4 * C), gr, br)
    y, mean, rstd = ops.ln_fwd(mode, x.to(DEV), gm.to(DEV), bt.to(DEV), B, H, W, C, ws, shift, 1e-5, code)
    tol = 1e-5 if ydt == "f32" else 6e-3
    assert rel(y.float().reshape(yr.shape), yr) < tol
    cot = torch.from_numpy(rng.standard_normal(tuple(yr.shape)).astype(np.float32))
    dres = torch.from_numpy(rng.standard_normal(tuple(x.shape)).astype(np.float32))
    cot_d = cot.to(DEV).to(y.dtype)
    (yr * cot_d.float().cpu().double()).sum().backward()
    dx, dg, db = ops.ln_bwd(mode, cot_d.reshape(y.shape).contiguous(), x.to(DEV), gm.to(DEV), mean, rstd, dres.to(DEV), B, H, W, C, ws, shift)
    assert rel(dx, xr.grad + dres.double()) < 1e-5
    assert rel(dg, gr.grad) < 1e-4
    assert rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("ydt", ["f32", "bf16"])
def test_ln_bwd_emits_partitioned_proj_dy(ydt):
    """ln_bwd (mode 0) with the fused second output == scale_cast(mode 1)(dx) + column sums."""
    ops, L = _ops()
    B, H, W, C, ws, shift = 2, 9, 13, 96, 7, 3
    td = torch.float32 if ydt == "f32" else torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn(B, H * W, C, generator=g).to(DEV)
    gm = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    bt = torch.randn(C, generator=g).to(DEV)
    dy = torch.randn(B * H * W, C, generator=g).to(td).to(DEV)
    dres = torch.randn(B, H * W, C, generator=g).to(DEV)
    s = torch.tensor([0.5, 1.25], device=DEV)
    _, mean, rstd = ops.ln_fwd(0, x, gm, bt, B, H, W, C, 1, 0, 1e-5, L.F32 if ydt == "f32" else L.BF16)
    dx0, dg0, db0 = ops.ln_bwd(0, dy, x, gm, mean, rstd, dres, B, H, W, C, 1, 0)
    for sh in (shift, 0):                       # shifted and unshifted window grids (pad slots must come back as exact zeros)
        junk = torch.full((B * 2 * 2 * 49 * C,), 7.0, device=DEV)       # dirty the allocator block dy2 is likely to reuse
        del junk
        dx1, dg1, db1, dy2, cs2 = ops.ln_bwd(0, dy, x, gm, mean, rstd, dres, B, H, W, C, 1, 0, emit_windows=(ws, sh, s))
        assert torch.equal(dx0, dx1) and rel(dg1, dg0) < 1e-5 and rel(db1, db0) < 1e-5
        want, wcs = ops.scale_cast(dx0, s, 1, B, H, W, C, ws, sh, L.F32 if ydt == "f32" else L.BF16, want_colsum=True)
        assert torch.equal(dy2, want)
        assert rel(cs2, wcs) < 1e-5


def test_scale_cast_and_colsum():
    ops, L = _ops()
    B, H, W, C, ws, shift = 2, 9, 13, 64, 7, 3
    x = torch.randn(B, H * W, C, device=DEV)
    s = torch.tensor([0.0, 1.25], device=DEV)
    y = ops.scale_cast(x, s, 0, B, H, W, C, 1, 0, L.BF16)
    assert torch.equal(y.view(B, H * W, C), (x * s.view(B, 1, 1)).bfloat16())
    yw, cs = ops.scale_cast(x, s, 1, B, H, W, C, ws, shift, L.F32, want_colsum=True)
    want = so.shift_gather((x * s.view(B, 1, 1)).cpu(), H, W, ws, shift)
    assert torch.equal(yw.cpu().view(want.shape), want)
    assert rel(cs, want.double().sum((0, 1))) < 1e-5                  # fused bias-gradient column sums
    y2, cs2 = ops.scale_cast(x, s, 0, B, H, W, C, 1, 0, L.BF16, want_colsum=True)
    assert torch.equal(y2, y) and rel(cs2, (x * s.view(B, 1, 1)).double().sum((0, 1))) < 1e-5
    for dt in (torch.float32, torch.bfloat16):
        X = torch.randn(1000, 96, device=DEV).to(dt)
        assert rel(ops.colsum(X), X.double().sum(0)) < 1e-5
    w = torch.randn(777, device=DEV)
    assert torch.equal(ops.cast_bf16(w), w.bfloat16())


def test_scale_cast_grad_gather():
    """Bucket packing of the data-parallel all-reduce: many ragged fp32 tensors -> one flat bucket, bit-exact; more tensors
    than one launch takes (GATHER_MAX), sizes that are not multiples of 4, a source that is only 4-byte aligned."""
    ops, L = _ops()
    g = torch.Generator(device="cpu").manual_seed(5)
    sizes = [1, 3, 4, 169 * 3, 96, 4097, 8192 + 5, 288 * 96, 7] + [int(x) for x in torch.randint(1, 5000, (L.GATHER_MAX + 9,), generator=g)]
    base = torch.randn(sum(sizes) + 8, generator=g).to(DEV)
    srcs, offs, off, pos = [], [], 0, 1            # pos starts at 1: the first source is not 16-byte aligned
    for n in sizes:
        srcs.append(base[pos:pos + n])
        pos += n
        offs.append(off)
        off += (n + 3) // 4 * 4
    bucket = torch.full((off,), -7.0, device=DEV)
    ops.grad_gather(srcs, offs, bucket)
    want = torch.full((off,), -7.0, device=DEV)
    for t, o in zip(srcs, offs):
        want[o:o + t.numel()] = t
    assert torch.equal(bucket, want)


# C <= 192: 32-token tiles with staged x rows; 192 < C <= 384: 32-token tiles, x from global memory; wider: 16-token tiles
@pytest.mark.parametrize("C,H,W", [(96, 9, 13), (768, 5, 7), (32, 40, 33), (192, 11, 9), (256, 6, 11), (384, 9, 13), (512, 7, 9),
                                   (1024, 5, 7), (384, 50, 84)])
def test_ln_nchw_fwd_bwd(C, H, W):
    ops, _ = _ops()
    B = 2
    rng = np.random.default_rng(9)
    x = torch.from_numpy(rng.standard_normal((B, H * W, C)).astype(np.float32))
    gm = torch.from_numpy((1 + 0.2 * rng.standard_normal(C)).astype(np.float32))
    bt = torch.from_numpy((0.3 * rng.standard_normal(C)).astype(np.float32))
    xr, gr, br = x.double().requires_grad_(True), gm.double().requires_grad_(True), bt.double().requires_grad_(True)
    yr = so.layer_norm(xr, gr, br).reshape(B, H, W, C).permute(0, 3, 1, 2)
    out, mean, rstd = ops.ln_nchw_fwd(x.to(DEV), gm.to(DEV), bt.to(DEV), H, W, 1e-5)
    assert out.shape == (B, C, H, W) and out.is_contiguous()
    assert rel(out, yr) < 1e-5
    cot = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32))
    (yr * cot.double()).sum().backward()
    dx, dg, db = ops.ln_nchw_bwd(cot.to(DEV), x.to(DEV), gm.to(DEV), mean, rstd)
    assert rel(dx, xr.grad) < 1e-5 and rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("Hi,Wi", [(50, 70), (32, 32), (17, 9)])
def test_patch_unfold_roundtrip_and_layout(dtype, Hi, Wi):
    ops, L = _ops()
    code = L.F32 if dtype == "f32" else L.BF16
    B, Cin, p = 2, 3, 4
    img = torch.randint(-100, 100, (B, Cin, Hi, Wi)).float()
    cols = ops.patch_gather(img.to(DEV), p, code)
    pad = torch.nn.functional.pad(img, (0, (-Wi) % p, 0, (-Hi) % p))
    Hh, Ww = pad.shape[2] // p, pad.shape[3] // p
    want = pad.reshape(B, Cin, Hh, p, Ww, p).permute(0, 2, 4, 1, 3, 5).reshape(B * Hh * Ww, Cin * p * p)
    assert torch.equal(cols.float().cpu(), want)                       # bit-exact (small integers are exact in bf16)
    back = ops.patch_scatter(cols, B, Cin, Hi, Wi, p)
    assert torch.equal(back.cpu(), img)


# ---------------------------------------------------------------- GEMM
def _gemm_ref(A, Bm, a_trans, b_trans):
    A2 = A.double().t() if a_trans else A.double()
    B2 = Bm.double() if b_trans else Bm.double().t()
    return A2 @ B2


GEMM_SHAPES = [(300, 96, 96), (257, 288, 96), (130, 384, 200), (128, 32, 64), (1000, 256, 512), (64, 768, 3072), (200, 48, 96), (96, 96, 48), (5000, 48, 48), (3000, 16, 64)]


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("a_trans,b_trans", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_store_all_layouts(dtype, a_trans, b_trans, M, N, K):
    ops, L = _ops()
    td = torch.float32 if dtype == "f32" else torch.bfloat16
    if dtype == "bf16" and a_trans and M % 8:
        pytest.skip("transposed bf16 operand needs a 16-byte row pitch")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn((K, M) if a_trans else (M, K), generator=g) * 0.5).to(td).to(DEV)
    Bm = (torch.randn((K, N) if b_trans else (N, K), generator=g) * 0.5).to(td).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    out = ops.gemm(A, Bm, M, N, K, a_trans=a_trans, b_trans=b_trans, bias=bias, out_dtype=L.F32)
    ref = _gemm_ref(A.cpu(), Bm.cpu(), a_trans, b_trans) + bias.cpu().double()
    assert rel(out, ref) < (1e-5 if dtype == "f32" else 1e-5), (M, N, K)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_gemm_epilogues(dtype):
    ops, L = _ops()
    td = torch.float32 if dtype == "f32" else torch.bfloat16
    tol_store = 1e-5 if dtype == "f32" else 6e-3
    B, H, W, C, ws, shift = 2, 9, 13, 64, 7, 3
    T = B * H * W
    g = torch.Generator(device="cpu").manual_seed(3)
    X = (torch.randn(T, C, generator=g)).to(td).to(DEV)
    Wt = (torch.randn(4 * C, C, generator=g) * 0.2).to(td).to(DEV)
    b = torch.randn(4 * C, generator=g).to(DEV)
    # GELU: D = gelu(u), D2 = gelu'(u)
    dg = torch.empty(T, 4 * C, dtype=td, device=DEV)
    h = ops.gemm(X, Wt, T, 4 * C, C, bias=b, epilogue=L.EPI_GELU, out2=dg)
    ur = (X.double().cpu() @ Wt.double().cpu().t() + b.double().cpu()).requires_grad_(True)
    hr = so.gelu_erf(ur)
    hr.sum().backward()
    assert rel(h, hr) < tol_store and rel(dg, ur.grad) < tol_store
    # DGELU: D = acc * aux (aux = saved gelu')
    dy = (torch.randn(T, C, generator=g)).to(td).to(DEV)
    du = ops.gemm(dy, Wt, T, 4 * C, C, b_trans=False, epilogue=L.EPI_DGELU, aux=dg)
    assert rel(du, (dy.double().cpu() @ Wt.double().cpu().t()) * dg.double().cpu()) < tol_store
    # RESIDUAL with per-image scale
    W2 = (torch.randn(C, 4 * C, generator=g) * 0.1).to(td).to(DEV)
    b2 = torch.randn(C, generator=g).to(DEV)
    res = torch.randn(T, C, generator=g).to(DEV)
    s = torch.tensor([0.0, 1.25], device=DEV)
    out = torch.empty(T, C, device=DEV)
    ops.gemm(h, W2, T, C, 4 * C, bias=b2, epilogue=L.EPI_RESIDUAL, out=out, aux=res, row_scale=s, rows_per_image=H * W)
    y = h.double().cpu() @ W2.double().cpu().t() + b2.double().cpu()
    ref = res.double().cpu() + (y.view(B, H * W, C) * s.double().cpu().view(B, 1, 1)).view(T, C)
    assert rel(out, ref) < 1e-5 if dtype == "f32" else rel(out, ref) < 2e-3
    # SCATTER_RESIDUAL: rows are window slots
    nslots = so.gather_index(H, W, ws, shift).shape[0]
    Ow = torch.randn(B * nslots, C, generator=g).to(td).to(DEV)
    Wp_ = (torch.randn(C, C, generator=g) * 0.2).to(td).to(DEV)
    out = torch.empty(T, C, device=DEV)
    ops.gemm(Ow, Wp_, B * nslots, C, C, bias=b2, epilogue=L.EPI_SCATTER_RESIDUAL, out=out, aux=res, row_scale=s, geom=(H, W, ws, shift))
    yw = (Ow.double().cpu() @ Wp_.double().cpu().t() + b2.double().cpu()).view(-1, ws * ws, C)
    ysc = so.shift_scatter(yw, B, H, W, ws, shift) * s.double().cpu().view(B, 1, 1)
    assert rel(out, res.double().cpu() + ysc.view(T, C)) < (1e-5 if dtype == "f32" else 2e-3)
    # ATOMIC_ADD split-K weight gradient: dW = dY^T X
    dW = torch.zeros(4 * C, C, device=DEV)
    db = torch.zeros(4 * C, device=DEV)
    dY = torch.randn(T, 4 * C, generator=g).to(td).to(DEV)
    ops.gemm(dY, X, 4 * C, C, T, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=dW, colsum_a=db)
    assert rel(dW, dY.double().cpu().t() @ X.double().cpu()) < 1e-5
    assert rel(db, dY.double().cpu().sum(0)) < 1e-5          # bias gradient from the all-ones MMA


def test_gemm_bf16_matches_fp32_kernel_on_device():
    """Cross-check of the two GEMM implementations on identical (bf16-representable) data, large split-K."""
    ops, L = _ops()
    T, Co, Ci = 50000, 288, 96
    dY = torch.randn(T, Co, device=DEV).bfloat16()
    X = torch.randn(T, Ci, device=DEV).bfloat16()
    d16 = torch.zeros(Co, Ci, device=DEV)
    d32 = torch.zeros(Co, Ci, device=DEV)
    b16 = torch.zeros(Co, device=DEV)
    b32 = torch.zeros(Co, device=DEV)
    ops.gemm(dY, X, Co, Ci, T, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=d16, colsum_a=b16)
    ops.gemm(dY.float(), X.float(), Co, Ci, T, a_trans=True, b_trans=True, epilogue=L.EPI_ATOMIC_ADD, out=d32, colsum_a=b32)
    assert rel(d16, d32) < 1e-4
    assert rel(b16, dY.double().sum(0)) < 1e-5 and rel(b32, dY.double().sum(0)) < 1e-5


# Both tile families of the tcgen05 GEMM on the same shapes: 1-CTA tiles (mode 0) and CTA-pair tiles forced wherever they
# are legal (mode 2; cta_group::2, each CTA stages half of B).  Shapes cover one pair tile, ragged / odd tile counts, half-tiles of
# 8..128 rows, MN-major operands, every fused epilogue, split-K with the tensor-core bias gradient, and multi-wave grids.
PAIR_CASES = [
    # name, M, N, K, epilogue, a_trans, b_trans, f32 output
    ("one unit", 256, 256, 64, "store", False, False, True),
    ("N=96 half 48", 256, 96, 192, "store", False, False, True),
    ("N=16 half 8", 256, 16, 128, "store", False, False, True),
    ("ragged M", 1000, 512, 320, "store", False, False, True),
    ("odd tiles, partial k-block", 1100, 192, 96, "store", False, False, True),
    ("bf16 TMA store N=192", 1024, 384, 384, "store", False, False, False),
    ("B MN-major", 1024, 512, 512, "store", False, True, True),
    ("B MN-major N=384", 1024, 384, 512, "store", False, True, False),
    ("A MN-major", 1024, 256, 512, "store", True, False, True),
    ("gelu", 1024, 512, 128, "gelu", False, False, False),
    ("dgelu", 1024, 512, 128, "dgelu", False, True, False),
    ("residual N=384", 2048, 384, 1536, "residual", False, False, True),
    ("split-K dW + colsum", 512, 256, 20000, "atomic", True, True, True),
    ("split-K dW N=384", 1536, 384, 16800, "atomic", True, True, True),
    ("many waves", 40000, 1152, 384, "store", False, False, False),
]


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("case", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
def test_gemm_tile_modes(case, mode):
    ops, L = _ops()
    _, M, N, K, epi, a_trans, b_trans, out_f32 = case
    prev = L.lib().swin_gemm_pair_mode(mode)
    try:
        g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
        A = (torch.randn((K, M) if a_trans else (M, K), generator=g) * 0.5).bfloat16().to(DEV)
        Bm = (torch.randn((K, N) if b_trans else (N, K), generator=g) * 0.5).bfloat16().to(DEV)
        acc = _gemm_ref(A.cpu(), Bm.cpu(), a_trans, b_trans)
        bias = torch.randn(N, generator=g).to(DEV)
        bd = bias.double().cpu()
        kw, want2 = {}, None
        if epi == "store":
            out = torch.empty(M, N, device=DEV, dtype=torch.float32 if out_f32 else torch.bfloat16)
            code, want = L.EPI_STORE, acc + bd
        elif epi == "gelu":
            out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
            kw["out2"] = torch.empty_like(out)
            u = (acc + bd).requires_grad_(True)
            hr = so.gelu_erf(u)
            hr.sum().backward()
            code, want, want2 = L.EPI_GELU, hr.detach(), u.grad
        elif epi == "dgelu":
            out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
            kw["aux"] = torch.randn(M, N, generator=g).bfloat16().to(DEV)
            code, want = L.EPI_DGELU, (acc + bd) * kw["aux"].double().cpu()
        elif epi == "residual":
            out = torch.empty(M, N, device=DEV, dtype=torch.float32)
            kw["aux"] = torch.randn(M, N, generator=g).to(DEV)
            code, want = L.EPI_RESIDUAL, kw["aux"].double().cpu() + acc + bd
        else:
            out = torch.zeros(M, N, device=DEV, dtype=torch.float32)
            kw["colsum_a"] = torch.zeros(M, device=DEV)
            code, want, bias = L.EPI_ATOMIC_ADD, acc, None
        ops.gemm(A, Bm, M, N, K, a_trans=a_trans, b_trans=b_trans, epilogue=code, bias=bias, out=out, **kw)
        tol = 1e-5 if out.dtype == torch.float32 else 6e-3       # fp32 accumulate of bf16 products / one bf16 output rounding
        assert rel(out, want) < tol, case[0]
        if want2 is not None:
            assert rel(kw["out2"], want2) < tol
        if "colsum_a" in kw:
            A2 = A.double().cpu().t() if a_trans else A.double().cpu()
            assert rel(kw["colsum_a"], A2.sum(1)) < 1e-5
    finally:
        L.lib().swin_gemm_pair_mode(prev)


def test_gemm_tile_modes_agree_bitwise_on_fp32_store():
    """The two tile families accumulate each output element over k in the same order (fp32, in TMEM): identical bits."""
    ops, L = _ops()
    M, N, K = 3000, 384, 768
    A = torch.randn(M, K, device=DEV).bfloat16()
    Bm = torch.randn(N, K, device=DEV).bfloat16()
    outs = []
    prev = L.lib().swin_gemm_pair_mode(-1)
    try:
        for mode in (0, 2):
            L.lib().swin_gemm_pair_mode(mode)
            outs.append(ops.gemm(A, Bm, M, N, K, out_dtype=L.F32).clone())
    finally:
        L.lib().swin_gemm_pair_mode(prev)
    assert torch.equal(outs[0], outs[1])


# ---------------------------------------------------------------- window attention core
def _attn_ref(qkv, bias, mask, nH, scale, cot):
    B_, N, C3 = qkv.shape
    C = C3 // 3
    q = qkv.double().requires_grad_(True)
    bz = bias.double().requires_grad_(True)
    qq = q[..., :C].reshape(B_, N, nH, 32).transpose(1, 2) * scale
    kk = q[..., C:2 * C].reshape(B_, N, nH, 32).transpose(1, 2)
    vv = q[..., 2 * C:].reshape(B_, N, nH, 32).transpose(1, 2)
    s = qq @ kk.transpose(-1, -2) + bz[None]
    if mask is not None:
        s = s + mask.double()[torch.arange(B_) % mask.shape[0]][:, None]
    p = torch.softmax(s, -1)
    o = (p @ vv).transpose(1, 2).reshape(B_, N, C)
    lse = torch.logsumexp(s, -1)
    (o * cot.double()).sum().backward()
    return o.detach(), lse.detach(), q.grad, bz.grad


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("B_,nH,masked", [(6, 2, True), (5, 3, False), (1, 1, False), (48, 6, True)])
def test_window_attention_core(dtype, B_, nH, masked):
    ops, L = _ops()
    td = torch.float32 if dtype == "f32" else torch.bfloat16
    ws, N, C = 7, 49, nH * 32
    g = torch.Generator(device="cpu").manual_seed(B_ * 10 + nH)
    qkv = torch.randn(B_, N, 3 * C, generator=g).to(td)
    bias = (torch.randn(nH, N, N, generator=g) * 0.5)
    mask = torch.from_numpy(so.shift_mask_np(10, 13, ws, 3)[1:4]) if masked else None   # nW = 3
    cot = torch.randn(B_, N, C, generator=g).to(td)
    o_r, lse_r, dqkv_r, dbias_r = _attn_ref(qkv.float(), bias, mask, nH, 32 ** -0.5, cot.float())
    md = mask.to(DEV) if masked else None
    o, lse = ops.window_attn_fwd(qkv.to(DEV), bias.to(DEV), md, B_, nH, ws, 32 ** -0.5)
    tol = 1e-5 if dtype == "f32" else 8e-3
    assert rel(o, o_r) < tol
    assert rel(lse, lse_r) < 1e-5 if dtype == "f32" else rel(lse, lse_r) < 2e-3
    dqkv, dbias = ops.window_attn_bwd(qkv.to(DEV), o, cot.to(DEV), lse, bias.to(DEV), md, B_, nH, ws, 32 ** -0.5)
    tolb = 1e-5 if dtype == "f32" else 1.5e-2
    C_ = C
    assert rel(dqkv[..., :C_], dqkv_r[..., :C_]) < tolb, "dQ"
    assert rel(dqkv[..., C_:2 * C_], dqkv_r[..., C_:2 * C_]) < tolb, "dK"
    assert rel(dqkv[..., 2 * C_:], dqkv_r[..., 2 * C_:]) < tolb, "dV"
    assert rel(dbias, dbias_r) < tolb, "dBias"


@pytest.mark.parametrize("H,W", [(10, 13), (25, 42), (7, 7)])
def test_window_attention_closed_form_canonical_mask_equals_tensor_mask(H, W):
    """bf16 kernel: evaluating the canonical SW-MSA mask in closed form == reading the (nW,49,49) tensor."""
    ops, L = _ops()
    ws, N, nH = 7, 49, 2
    nwh, nww = -(-H // ws), -(-W // ws)
    nW = nwh * nww
    B_ = 3 * nW
    g = torch.Generator(device="cpu").manual_seed(5)
    qkv = torch.randn(B_, N, 3 * nH * 32, generator=g).bfloat16().to(DEV)
    bias = (torch.randn(nH, N, N, generator=g) * 0.5).to(DEV)
    dout = torch.randn(B_, N, nH * 32, generator=g).bfloat16().to(DEV)
    mask = ops.shift_mask(H, W, ws, 3, DEV)
    assert torch.equal(mask.cpu(), torch.from_numpy(so.shift_mask_np(H, W, ws, 3)))
    nz = ops.mask_nonzero(mask)
    o1, l1 = ops.window_attn_fwd(qkv, bias, mask, B_, nH, ws, 32 ** -0.5, nz)
    o2, l2 = ops.window_attn_fwd(qkv, bias, mask, B_, nH, ws, 32 ** -0.5, nz, canon=(nwh, nww))
    assert torch.equal(o1, o2) and torch.equal(l1, l2)
    d1, b1 = ops.window_attn_bwd(qkv, o1, dout, l1, bias, mask, B_, nH, ws, 32 ** -0.5, nz)
    d2, b2 = ops.window_attn_bwd(qkv, o1, dout, l1, bias, mask, B_, nH, ws, 32 ** -0.5, nz, canon=(nwh, nww))
    assert torch.equal(d1, d2) and rel(b2, b1) < 1e-5


@pytest.mark.parametrize("B_,nH,grid", [(16 * 1392, 3, (200, 334)), (16 * 1392 - 1, 3, None), (16 * 96, 12, (50, 84))])
def test_window_attention_full_size_tcgen05_vs_fp32_kernels(B_, nH, grid):
    """BASELINE full sizes (Swin-T stage 0 / stage 2 at B=16, 800x1333): the persistent tcgen05 kernels (every CTA walks
    ~100 items; odd window count exercises the half-empty last tile) against the fp32 FFMA kernels on the same data."""
    ops, L = _ops()
    ws, N, C = 7, 49, nH * 32
    g = torch.Generator(device=DEV).manual_seed(B_ + nH)
    qkv = torch.randn(B_, N, 3 * C, generator=g, device=DEV)
    bias = torch.randn(nH, N, N, generator=g, device=DEV) * 0.5
    dout = torch.randn(B_, N, C, generator=g, device=DEV)
    mask = nz = None
    canon = (0, 0)
    if grid is not None:
        mask = ops.shift_mask(grid[0], grid[1], ws, 3, DEV)
        nz = ops.mask_nonzero(mask)
        canon = (-(-grid[0] // ws), -(-grid[1] // ws))
    o32, l32 = ops.window_attn_fwd(qkv, bias, mask, B_, nH, ws, 32 ** -0.5, nz)
    d32, b32 = ops.window_attn_bwd(qkv, o32, dout, l32, bias, mask, B_, nH, ws, 32 ** -0.5, nz)
    q16, do16 = qkv.bfloat16(), dout.bfloat16()
    o16, l16 = ops.window_attn_fwd(q16, bias, mask, B_, nH, ws, 32 ** -0.5, nz, canon)
    assert rel(o16, o32) < 8e-3 and rel(l16, l32) < 2e-3
    assert torch.isfinite(o16.float()).all()
    d16, b16 = ops.window_attn_bwd(q16, o16, do16, l16, bias, mask, B_, nH, ws, 32 ** -0.5, nz, canon)
    assert rel(d16, d32) < 1.5e-2 and rel(b16, b32) < 1.5e-2


@pytest.mark.parametrize("B_,nH,masked", [(6, 3, True), (1, 1, False), (301, 4, False), (600, 2, True)])
def test_window_attention_core_window12_bf16(B_, nH, masked):
    """Window 12 in bf16 mode (the 384-pixel Swin-B/L configs): the warp-level tensor-core kernels (attn_mma.cu, bf16 q/k/v/out/
    dout/dqkv, fp32 softmax statistics) against the fp32 core on the same (bf16-rounded) inputs, forward and backward."""
    ops, L = _ops()
    ws, N = 12, 144
    g = torch.Generator(device="cpu").manual_seed(13 + B_)
    qkv = torch.randn(B_, N, 3 * nH * 32, generator=g).bfloat16()
    bias = (torch.randn(nH, N, N, generator=g) * 0.5).to(DEV)
    mask = nz = None
    if masked:
        mask = torch.from_numpy(so.shift_mask_np(24, 36, ws, 6)).to(DEV)           # nW = 6
        nz = ops.mask_nonzero(mask)
    cot = torch.randn(B_, N, nH * 32, generator=g).bfloat16()
    o32, lse32 = ops.window_attn_fwd(qkv.float().to(DEV), bias, mask, B_, nH, ws, 32 ** -0.5)
    d32, b32 = ops.window_attn_bwd(qkv.float().to(DEV), o32, cot.float().to(DEV), lse32, bias, mask, B_, nH, ws, 32 ** -0.5)
    o16, lse16 = ops.window_attn_fwd(qkv.to(DEV), bias, mask, B_, nH, ws, 32 ** -0.5, nz)
    assert o16.dtype == torch.bfloat16 and rel(o16, o32) < 6e-3, rel(o16, o32)
    assert torch.allclose(lse16, lse32, rtol=1e-4, atol=1e-4)
    d16, b16 = ops.window_attn_bwd(qkv.to(DEV), o16, cot.to(DEV), lse16, bias, mask, B_, nH, ws, 32 ** -0.5, nz)
    assert d16.dtype == torch.bfloat16 and rel(d16, d32) < 1.5e-2, rel(d16, d32)
    assert rel(b16, b32) < 1.5e-2, rel(b16, b32)


def test_window_attention_core_window12_fp32():
    """BASELINE config 5 also sweeps window 12: served by the fp32 kernels (forward and backward)."""
    ops, L = _ops()
    ws, N, nH, B_ = 12, 144, 2, 4
    g = torch.Generator(device="cpu").manual_seed(12)
    qkv = torch.randn(B_, N, 3 * nH * 32, generator=g)
    bias = torch.randn(nH, N, N, generator=g) * 0.5
    mask = torch.from_numpy(so.shift_mask_np(20, 33, ws, 6)[2:6])          # nW = 4
    cot = torch.randn(B_, N, nH * 32, generator=g)
    o_r, lse_r, dqkv_r, dbias_r = _attn_ref(qkv, bias, mask, nH, 32 ** -0.5, cot)
    o, lse = ops.window_attn_fwd(qkv.to(DEV), bias.to(DEV), mask.to(DEV), B_, nH, ws, 32 ** -0.5)
    assert rel(o, o_r) < 1e-5 and rel(lse, lse_r) < 1e-5
    dqkv, dbias = ops.window_attn_bwd(qkv.to(DEV), o, cot.to(DEV), lse, bias.to(DEV), mask.to(DEV), B_, nH, ws, 32 ** -0.5)
    assert rel(dqkv, dqkv_r) < 1e-5 and rel(dbias, dbias_r) < 1e-5
