"""ctypes binding of the C ABI declared in include/swin_b200.h (libswin_b200.so).

There is deliberately no fallback: if the shared library is missing or a call fails, a
RuntimeError is raised (the product must not silently run anything but the sm_100a kernels).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SWIN_B200_LIB") or os.path.join(_HERE, "libswin_b200.so")   # env override: development builds only
CSRC = os.path.join(_HERE, "csrc")

F32, BF16 = 0, 1
GATHER_MAX = 64
EPI_STORE, EPI_GELU, EPI_RESIDUAL, EPI_SCATTER_RESIDUAL, EPI_DGELU, EPI_ATOMIC_ADD = range(6)

c_int, c_i64, c_f32, vp = C.c_int, C.c_int64, C.c_float, C.c_void_p


class LnArgs(C.Structure):
    _fields_ = [("mode", c_int), ("B", c_int), ("H", c_int), ("W", c_int), ("C", c_int), ("ws", c_int), ("shift", c_int),
                ("eps", c_f32), ("y_dtype", c_int), ("x", vp), ("gamma", vp), ("beta", vp), ("y", vp), ("mean", vp),
                ("rstd", vp), ("dy", vp), ("dres", vp), ("dx", vp), ("dgamma", vp), ("dbeta", vp),
                ("dy2", vp), ("dy2_scale", vp), ("dy2_colsum", vp), ("ws2", c_int), ("shift2", c_int)]


class GemmArgs(C.Structure):
    _fields_ = [("dtype", c_int), ("M", c_int), ("N", c_int), ("K", c_int),
                ("A", vp), ("a_trans", c_int), ("lda", c_i64),
                ("B", vp), ("b_trans", c_int), ("ldb", c_i64),
                ("epilogue", c_int), ("bias", vp),
                ("D", vp), ("d_dtype", c_int), ("ldd", c_i64),
                ("D2", vp), ("aux", vp), ("row_scale", vp), ("rows_per_image", c_int),
                ("H", c_int), ("W", c_int), ("ws", c_int), ("shift", c_int), ("colsum_a", vp)]


class AttnArgs(C.Structure):
    _fields_ = [("dtype", c_int), ("B_", c_int), ("nH", c_int), ("ws", c_int), ("nW", c_int), ("scale", c_f32),
                ("qkv", vp), ("bias", vp), ("mask", vp), ("mask_nz", vp), ("canon_nwh", c_int), ("canon_nww", c_int), ("out", vp), ("lse", vp),
                ("dout", vp), ("dqkv", vp), ("dbias", vp)]


class AttnQkvArgs(C.Structure):
    _fields_ = [("B_", c_int), ("nH", c_int), ("ws", c_int), ("nW", c_int), ("scale", c_f32), ("x", vp), ("wqkv", vp), ("bqkv", vp),
                ("bias", vp), ("mask", vp), ("mask_nz", vp), ("canon_nwh", c_int), ("canon_nww", c_int), ("out", vp), ("lse", vp),
                ("qkv_out", vp), ("workspace", vp), ("workspace_bytes", C.c_longlong)]


# name -> (restype, argtypes): every symbol include/swin_b200.h declares
SYMBOLS = {
    "swin_version": (c_int, []),
    "swin_last_error": (C.c_char_p, []),
    "swin_device_check": (c_int, [c_int]),
    "swin_sm_reserve": (c_int, [c_int]),
    "swin_window_partition": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, vp]),
    "swin_window_reverse": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, vp]),
    "swin_window_gather": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, vp]),
    "swin_window_scatter": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, vp]),
    "swin_shift_mask": (c_int, [vp, c_int, c_int, c_int, c_int, vp]),
    "swin_mask_nonzero": (c_int, [vp, vp, c_int, c_int, vp]),
    "swin_rel_bias_expand": (c_int, [vp, vp, c_int, c_int, vp]),
    "swin_rel_bias_reduce": (c_int, [vp, vp, c_int, c_int, vp]),
    "swin_ln_fwd": (c_int, [C.POINTER(LnArgs), vp]),
    "swin_ln_bwd": (c_int, [C.POINTER(LnArgs), vp]),
    "swin_ln_nchw_fwd": (c_int, [vp, vp, vp, vp, vp, vp, c_int, c_int, c_int, c_f32, vp]),
    "swin_ln_nchw_bwd": (c_int, [vp, vp, vp, vp, vp, vp, vp, vp, c_int, c_int, c_int, vp]),
    "swin_patch_gather": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, vp]),
    "swin_patch_scatter": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, vp]),
    "swin_gemm": (c_int, [C.POINTER(GemmArgs), vp]),
    "swin_colsum": (c_int, [vp, c_int, c_int, c_i64, c_int, vp, vp]),
    "swin_scale_cast": (c_int, [vp, vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, vp, vp]),
    "swin_cast_bf16": (c_int, [vp, vp, c_i64, vp]),
    "swin_gemm_pair_mode": (c_int, [c_int]),
    "swin_gemm_plan": (c_int, [vp, vp]),
    "swin_grad_gather": (c_int, [vp, vp, vp, c_int, vp, vp]),
    "swin_adamw_step": (c_int, [vp, vp, vp, vp, vp, vp, vp, c_int, C.c_double, C.c_double, C.c_double, C.c_double, c_int, C.c_double, vp]),
    "swin_window_attn_fwd": (c_int, [C.POINTER(AttnArgs), vp]),
    "swin_window_attn_bwd": (c_int, [C.POINTER(AttnArgs), vp]),
    "swin_window_attn_qkv_fwd": (c_int, [C.POINTER(AttnQkvArgs), vp]),
    "swin_window_attn_qkv_supported": (c_int, [c_int, c_int, c_int]),
    "swin_window_attn_qkv_workspace": (C.c_longlong, [c_int, c_int, c_int]),
}

_lib = None
_lock = threading.Lock()


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libswin_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libswin_b200.so failed")
    return LIB_PATH


def lib():
    """The loaded library; raises RuntimeError (no fallback) if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.isfile(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `make -C swin-transformer-object-detection_b200/csrc` — swin_b200 has no CPU/PyTorch fallback")
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in SYMBOLS.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().swin_last_error()
        raise RuntimeError(f"swin_b200 {what} failed (code {rc}): {msg.decode() if msg else ''}")
