"""ctypes binding of libo2b200.so (the C ABI declared in include/o2b200.h).

The product path fails loudly if the shared library is missing or no sm_100 device is present:
there is no CPU or PyTorch fallback behind any op.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("O2B200_LIB") or os.path.join(HERE, "libo2b200.so")   # override: A/B kernel experiments

O2_F32, O2_BF16 = 0, 1
GEMM_SIMT_F32, GEMM_TC_BF16 = 0, 1
EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_DGELU, EPI_ACCUM = range(6)
LOSS_MSE, LOSS_MAE, LOSS_BAYESIAN_TV = 0, 1, 2

_p, _i, _l, _f, _u = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32

SIGNATURES = {
    "o2_version": ([], C.c_int),
    "o2_last_error": ([], C.c_char_p),
    "o2_device_ok": ([], C.c_int),
    "o2_gemm": ([_i, _p, _i, _l, _p, _i, _l, _p, _i, _l, _l, _l, _l, _i, _p, _p, _l, _l, _p, _l, _i, _p], _i),
    "o2_gemm_drop": ([_p, _i, _l, _p, _i, _l, _p, _l, _l, _l, _l, _i, _p, _p, _l, _l, _p, _l, _p, _p], _i),
    "o2_layernorm_fwd": ([_p, _p, _p, _p, _p, _p, _l, _i, _f, _i, _p], _i),
    "o2_layernorm_bwd": ([_p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _i, _i, _p], _i),
    "o2_layernorm_bwd_drop": ([_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _i, _p, _p], _i),
    "o2_attn_fwd": ([_i, _p, _p, _p, _i, _i, _i, _i, _f, _p], _i),
    "o2_attn_bwd": ([_i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p], _i),
    "o2_attn_bwd_parts": ([_i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p], _i),
    "o2_attn_fwd_drop": ([_i, _p, _p, _p, _i, _i, _i, _i, _f, _f, C.c_uint64, _u, _p], _i),
    "o2_attn_bwd_parts_drop": ([_i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, C.c_uint64, _u, _p], _i),
    "o2_attn_bwd_fused_workspace": ([_i, _i, _i, _i], C.c_size_t),
    "o2_attn_bwd_fused": ([_i, _p, _p, _p, _p, _p, _p, _p, C.c_size_t, _i, _i, _i, _i, _f, _f, C.c_uint64, _u, _p], _i),
    "o2_frontend_fwd": ([_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "o2_frontend_bwd": ([_p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "o2_path2_conv1_fwd": ([_p, C.POINTER(C.c_int), _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "o2_path2_conv1_bwd": ([_p, C.POINTER(C.c_int), _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "o2_headtail_fwd": ([_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "o2_headtail_bwd": ([_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p], _i),
    "o2_loss_fwd_bwd": ([_p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _u, _i, _i, _i, _i, _i, _i, _f, _p], _i),
    "o2_clip_replace": ([_p, _i, _p, _i, _u, _i, _i, _i, _i, _i, _i, _p], _i),
    "o2_scale_channels": ([_p, _i, _p, _i, _i, _l, _p], _i),
    "o2_dropout": ([_p, _p, _p, _i, _l, _l, _l, _f, _p, C.c_uint64, _u, _p], _i),
    "o2_dropout_seed_source": ([_p], _i),
    "o2_normalize_fields": ([_p, _p, _p, _p, _i, _i, _l, _p], _i),
    "o2_eval_stats": ([_p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p], _i),
    "o2_cast_f32_to_bf16": ([_p, _p, _l, _p], _i),
    "o2_cast_bf16_to_f32": ([_p, _p, _l, _p], _i),
    "o2_bicubic_fwd": ([_p, _p, _i, _i, _i, _i, _i, _p], _i),
    "o2_bicubic_bwd": ([_p, _p, _i, _i, _i, _i, _i, _p], _i),
    "o2_colsum": ([_p, _i, _p, _l, _l, _l, _p], _i),
    "o2_adamw": ([_p, _p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _i, _f, _p], _i),
    "o2_adamw_dev": ([_p, _p, _p, _p, _p, _l, _p, _p], _i),
    "o2_nonfinite": ([_p, _l, _p, _p], _i),
}



class GemmDrop(C.Structure):
    """O2GemmDrop of include/o2b200.h"""
    _fields_ = [("p", C.c_float), ("seed", C.c_uint64), ("site", C.c_uint32), ("sample_scale", C.c_void_p),
                ("rows_per_sample", C.c_int64), ("after_residual", C.c_int32)]


_lib = None


class O2Error(RuntimeError):
    pass


def load(require_device: bool = False):
    """dlopen the library (no CUDA call is made unless require_device)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise O2Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(orbit2_b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(lib, name, None)     # tests/test_abi.py asserts every declared symbol exists
            if fn is None:
                continue
            fn.argtypes = args
            fn.restype = res
        _lib = lib
    if require_device and not _lib.o2_device_ok():
        raise O2Error("orbit2_b200 needs an sm_100a GPU: " + _lib.o2_last_error().decode())
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise O2Error(f"{what} failed ({rc}): {_lib.o2_last_error().decode()}")
