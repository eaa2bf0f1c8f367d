"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, argument validation returns errno-style codes without touching a GPU, the Python mirror keeps the
reference's constructor / state_dict contract, and there is no CPU fallback."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def lib():
    from swin_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "swin_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swin_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from swin_b200 import _lib
    names = header_symbols()
    assert len(names) >= 18
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (swin_[a-z0-9_]+)", nm))
    assert set(names) <= exported
    assert lib.swin_version() == 100


def test_only_sm100a_code_is_embedded():
    from swin_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_validation_is_errno_style_and_does_not_need_a_gpu(lib):
    from swin_b200 import _lib as L
    a = L.GemmArgs(dtype=L.F32, M=4, N=12, K=4, A=16, B=16, D=16, d_dtype=L.F32, ldd=12, lda=4, ldb=4)
    rc = lib.swin_gemm(ctypes.byref(a), None)
    assert rc == -22 and b"multiple of 8" in lib.swin_last_error()          # -EINVAL
    assert lib.swin_window_gather(16, 16, 1, 7, 7, 3, 7, 0, 2, None) == -22  # row bytes not a multiple of 16
    assert lib.swin_window_gather(16, 16, 1, 7, 7, 8, 7, 7, 2, None) == -22  # shift must be < ws
    assert lib.swin_shift_mask(16, 0, 7, 7, 3, None) == -22
    at = L.AttnArgs(dtype=L.BF16, B_=5, nH=3, ws=7, nW=2, qkv=16, bias=16, mask=16, out=16, lse=16, scale=1.0)
    assert lib.swin_window_attn_fwd(ctypes.byref(at), None) == -22
    assert b"multiple of nW" in lib.swin_last_error()
    aq = L.AttnQkvArgs(B_=4, nH=3, ws=12, x=16, wqkv=16, bias=16, out=16, scale=1.0)
    assert lib.swin_window_attn_qkv_fwd(ctypes.byref(aq), None) == -22
    assert b"window_size 7" in lib.swin_last_error()
    assert lib.swin_window_attn_qkv_supported(384, 12, 7) == 1 and lib.swin_window_attn_qkv_supported(512, 16, 7) == 0
    assert lib.swin_window_attn_qkv_workspace(96, 3, 7) == 3 * 52 * 64 * 4
    with pytest.raises(RuntimeError):
        L.check(-22, "demo")


def test_state_dict_contract_matches_reference():
    import swin_b200
    d = np.load(os.path.join(GOLDEN, "backbone_tiny.npz"))
    net = swin_b200.SwinTransformer()
    sd = net.state_dict()
    assert list(sd.keys()) == list(d["swin_t_keys"])
    assert [",".join(map(str, v.shape)) for v in sd.values()] == list(d["swin_t_shapes"])
    assert sd["layers.0.blocks.0.attn.relative_position_index"].dtype == torch.int64
    from oracle import swin_oracle as so
    assert torch.equal(sd["layers.0.blocks.0.attn.relative_position_index"], torch.from_numpy(so.relative_position_index_np(7)))
    assert sum(p.numel() for p in net.parameters()) == 27520698


def test_constructor_contract_and_schedule():
    import swin_b200
    net = swin_b200.SwinTransformer(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], drop_path_rate=0.3,
                                    out_indices=(1, 3), frozen_stages=2)
    rates = [b.drop_path.drop_prob if isinstance(b.drop_path, swin_b200.DropPath) else 0.0 for l in net.layers for b in l.blocks]
    assert np.allclose(rates, torch.linspace(0, 0.3, 24).numpy())
    assert [b.shift_size for b in net.layers[2].blocks][:4] == [0, 3, 0, 3]
    assert hasattr(net, "norm1") and hasattr(net, "norm3") and not hasattr(net, "norm0")
    # frozen_stages=2 freezes patch_embed and layer 0 and keeps them in eval after .train()
    net.train()
    assert not net.patch_embed.proj.weight.requires_grad and not net.layers[0].blocks[0].norm1.weight.requires_grad
    assert net.layers[1].blocks[0].norm1.weight.requires_grad
    assert not net.patch_embed.training and not net.layers[0].training and net.layers[1].training
    with pytest.raises(TypeError):
        net.init_weights(pretrained=123)


def test_paramwise_optimizer_keys_exist():
    """configs/swin/*: weight decay 0 for names containing these substrings -> they must exist under the same names."""
    import swin_b200
    names = [n for n, _ in swin_b200.SwinTransformer(ape=True).named_parameters()]
    for sub in ("absolute_pos_embed", "relative_position_bias_table", "norm"):
        assert any(sub in n for n in names)


def test_registry_contract():
    import swin_b200
    cfg = dict(type="SwinTransformer", embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7,
               ape=False, drop_path_rate=0.1, patch_norm=True, use_checkpoint=False)
    net = swin_b200.build_backbone(cfg)
    assert type(net).__name__ == "SwinTransformer" and cfg["type"] == "SwinTransformer"
    with pytest.raises(KeyError):
        swin_b200.build_backbone(dict(type="NoSuchBackbone"))


def test_no_cpu_fallback():
    import swin_b200
    net = swin_b200.SwinTransformer(embed_dim=32, depths=[2], num_heads=[1], out_indices=(0,))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        swin_b200.window_partition(torch.zeros(1, 7, 7, 8), 7)
    blk = swin_b200.SwinTransformerBlock(32, 1)
    blk.H = blk.W = 7
    with pytest.raises(RuntimeError):
        blk(torch.zeros(1, 49, 32), None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "swin-transformer-object-detection_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("oracle/", ""), f


def test_drop_path_scale_semantics():
    import swin_b200
    dp = swin_b200.DropPath(0.25)
    dp.train()
    torch.manual_seed(0)
    s = dp.sample_scale(torch.zeros(1000, 4, 8))
    assert s.shape == (1000,) and all(min(abs(v), abs(v - 1 / 0.75)) < 1e-6 for v in s.unique().tolist())
    assert abs((s > 0).float().mean().item() - 0.75) < 0.05
    dp.eval()
    assert dp.sample_scale(torch.zeros(3, 4, 8)) is None
    torch.manual_seed(0)
    dp.train()
    r = 0.75 + torch.rand((1000, 1, 1))
    torch.manual_seed(0)
    assert torch.equal(dp.sample_scale(torch.zeros(1000, 4, 8)), (r.floor() / 0.75).reshape(-1))


def test_gemm_pair_mode_knob_is_a_plain_setter(lib):
    """swin_gemm_pair_mode: returns the previous mode, a negative argument only queries, values above 2 clamp."""
    prev = lib.swin_gemm_pair_mode(-1)
    try:
        assert lib.swin_gemm_pair_mode(0) == prev
        assert lib.swin_gemm_pair_mode(-1) == 0
        assert lib.swin_gemm_pair_mode(7) == 0
        assert lib.swin_gemm_pair_mode(-1) == 2
    finally:
        lib.swin_gemm_pair_mode(prev)
    assert lib.swin_gemm_pair_mode(-1) == prev


def test_block_link_hands_over_only_the_unmodified_gradient():
    """functional.BlockLink: the next block's emitted dY is taken only for the very tensor it was derived from."""
    from swin_b200.functional import BlockLink
    dx = torch.zeros(2, 6, 8)
    dy2, cs = torch.ones(12, 8), torch.ones(8)
    link = BlockLink()
    assert link.take(dx) is None                               # nothing deposited
    link.deposit(dx, dy2, cs)
    got = link.take(dx.detach())                               # autograd hands over a detached alias: same storage, same version
    assert got is not None and got[0] is dy2 and got[1] is cs
    assert link.take(dx) is None                               # consumed: a second backward falls back to scale_cast
    link.deposit(dx, dy2, cs)
    assert link.take(dx.clone()) is None                       # another tensor (accumulated / hooked gradient)
    link.deposit(dx, dy2, cs)
    dx.add_(1.0)                                               # modified in place after the emission
    assert link.take(dx) is None
    link.deposit(dx, dy2, cs)
    assert link.take(dx.view(12, 8)) is None                   # same storage, different shape


def _plan(lib, M, N, K, epi=0, a_trans=0, b_trans=0):
    from swin_b200 import _lib as L
    a = L.GemmArgs(dtype=L.BF16, M=M, N=N, K=K, a_trans=a_trans, b_trans=b_trans, epilogue=epi)
    out = (ctypes.c_int * 6)()
    rc = lib.swin_gemm_plan(ctypes.byref(a), ctypes.cast(out, ctypes.c_void_p))
    assert rc == 0, lib.swin_last_error()
    return tuple(out)          # CTAs per tile, tile width, tiles along M, tiles along N, split-K factor, k-blocks per split


def test_gemm_tile_policy_on_the_benchmark_shapes(lib):
    """Host logic of the tcgen05 GEMM (no GPU needed): which tile family / width / split-K factor each GEMM of the
    benchmark step (Swin-T, B=16, 800x1333) gets.  The expected values are the ones the measurements in
    profiles/r01/gemm_pair_ab.txt were taken with."""
    from swin_b200 import _lib as L
    T0, Tp0, T2, Tp2 = 1068800, 1091328, 67200, 75264
    prev = lib.swin_gemm_pair_mode(1)
    try:
        # forward, K-major B: CTA pairs, each CTA stages half of the tile width
        assert _plan(lib, Tp2, 1152, 384) == (2, 192, 294, 6, 1, 6)                        # qkv, stage 2
        assert _plan(lib, T2, 1536, 384, L.EPI_GELU) == (2, 256, 263, 6, 1, 6)             # fc1 + GELU (odd tile count, >= 8 tiles)
        assert _plan(lib, T2, 384, 1536, L.EPI_RESIDUAL) == (2, 192, 263, 2, 1, 24)        # fc2 + residual
        assert _plan(lib, Tp2, 384, 384, L.EPI_SCATTER_RESIDUAL) == (2, 192, 294, 2, 1, 6) # proj + window reverse + residual
        assert _plan(lib, Tp0, 288, 96) == (2, 96, 4263, 3, 1, 2)                          # qkv, stage 0 (K = 96: 64 + 32)
        assert _plan(lib, T0, 96, 48) == (2, 96, 4175, 1, 1, 1)                            # PatchEmbed conv as a GEMM
        # backward: DGELU stays on 1-CTA tiles; MN-major B with N = 384 would need 128-wide pair tiles -> 1-CTA 192-wide
        assert _plan(lib, T2, 1536, 384, L.EPI_DGELU, b_trans=1) == (1, 256, 525, 6, 1, 6)
        assert _plan(lib, T2, 384, 1536, b_trans=1) == (1, 192, 525, 2, 1, 24)             # dX of fc1
        assert _plan(lib, 16800, 768, 3072, b_trans=1) == (2, 256, 66, 3, 1, 48)           # dX of fc1, stage 3 (N % 256 == 0)
        # split-K weight gradients: pairs only for an even tile count; the split fills the SMs (or SM pairs) in whole rounds
        assert _plan(lib, 1536, 384, T2, L.EPI_ATOMIC_ADD, 1, 1) == (2, 128, 6, 3, 4, 263)     # 18 tiles x 4 = 72 of 74 pairs
        assert _plan(lib, 1152, 384, Tp2, L.EPI_ATOMIC_ADD, 1, 1) == (1, 192, 9, 2, 8, 147)    # 9 tiles along M: odd -> 1-CTA, 144 of 148 SMs
        assert _plan(lib, 384, 1536, T2, L.EPI_ATOMIC_ADD, 1, 1) == (1, 256, 3, 6, 8, 132)
        # a single 128-row tile never pairs
        assert _plan(lib, 96, 384, T0, L.EPI_ATOMIC_ADD, 1, 1)[0] == 1
        lib.swin_gemm_pair_mode(0)
        assert _plan(lib, Tp2, 1152, 384) == (1, 192, 588, 6, 1, 6)
        lib.swin_gemm_pair_mode(2)                                                          # forced wherever legal (tests, A/B tools)
        assert _plan(lib, T2, 1536, 384, L.EPI_DGELU, b_trans=1) == (2, 256, 263, 6, 1, 6)
        assert _plan(lib, T2, 384, 1536, b_trans=1) == (2, 128, 263, 3, 1, 24)
        assert _plan(lib, T0, 96, 384, b_trans=1)[0] == 1                                  # N = 96, MN-major: no legal half-tile
    finally:
        lib.swin_gemm_pair_mode(prev)
    a = L.GemmArgs(dtype=L.F32, M=128, N=128, K=64)
    assert lib.swin_gemm_plan(ctypes.byref(a), ctypes.cast((ctypes.c_int * 6)(), ctypes.c_void_p)) == -22
