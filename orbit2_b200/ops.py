"""Thin tensor -> pointer wrappers over the C ABI.  Every function launches on torch's current stream
and raises O2Error on failure.  No op has a PyTorch/CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from ._lib import (EPI_ACCUM, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_DGELU, EPI_NONE, GEMM_SIMT_F32, GEMM_TC_BF16,
                   O2_BF16, O2_F32)

LAUNCHES = 0          # number of library kernels-launching calls (bench.py reports it)


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda, "orbit2_b200 ops need CUDA tensors (no CPU fallback)"
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return O2_F32
    if t.dtype == torch.bfloat16:
        return O2_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def impl_for(dtype: torch.dtype) -> int:
    return GEMM_TC_BF16 if dtype == torch.bfloat16 else GEMM_SIMT_F32


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, trans_a=False, trans_b=False, epi=EPI_NONE,
         bias=None, aux=None, aux_rows=0, aux_out=None, split_k=1, M=None, N=None, K=None):
    """out[M,N] = op(a) @ op(b) with fused epilogue; a/b/out are 2-D row-major (last stride 1)."""
    lib = L.load()
    assert a.dim() == 2 and b.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    if M is None:
        M = a.shape[1] if trans_a else a.shape[0]
    if K is None:
        K = a.shape[0] if trans_a else a.shape[1]
    if N is None:
        N = b.shape[1] if trans_b else b.shape[0]
    assert (b.shape[0] if trans_b else b.shape[1]) == K, (a.shape, b.shape, trans_a, trans_b)
    assert out.shape[0] == M and out.shape[1] == N
    impl = impl_for(a.dtype)
    assert b.dtype == a.dtype
    rc = lib.o2_gemm(impl, _ptr(a), int(trans_a), a.stride(0), _ptr(b), int(trans_b), b.stride(0), _ptr(out), dt(out),
                     out.stride(0), M, N, K, epi, _ptr(bias), _ptr(aux), aux.stride(0) if aux is not None else 0,
                     aux_rows, _ptr(aux_out), aux_out.stride(0) if aux_out is not None else 0, split_k, _stream())
    L.check(rc, "o2_gemm")
    _count()
    return out
