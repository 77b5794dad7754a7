"""Micro-benchmark of o2_gemm (tcgen05) at the 117M training shapes: TFLOP/s per shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from orbit2_b200 import ops  # noqa: E402
from orbit2_b200.reslim import _wgrad_split  # noqa: E402
from orbit2_b200._lib import EPI_ACCUM, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_DGELU, EPI_NONE  # noqa: E402

T, D, H = 129600, 1024, 4096
dev = "cuda"
bf = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: (torch.randn(*s, generator=g, device=dev) * 0.1).to(bf)
x, xh = rn(T, D), rn(T, H)
wqkv, wfc1, wfc2, wp = rn(3 * D, D), rn(H, D), rn(D, H), rn(D, D)
b3, b1, bh = torch.zeros(3 * D, device=dev), torch.zeros(D, device=dev), torch.zeros(H, device=dev)
dy3 = rn(T, 3 * D)
cases = [
    ("qkv  fwd  [T,D]x[3D,D]^T +bias", lambda: ops.gemm(x, wqkv, torch.empty(T, 3 * D, device=dev, dtype=bf), epi=EPI_BIAS, bias=b3), 2.0 * T * D * 3 * D),
    ("proj fwd  +bias+res", lambda: ops.gemm(x, wp, torch.empty(T, D, device=dev, dtype=bf), epi=EPI_BIAS_RES, bias=b1, aux=x), 2.0 * T * D * D),
    ("fc1  fwd  +bias+gelu", lambda: ops.gemm(x, wfc1, torch.empty(T, H, device=dev, dtype=bf), epi=EPI_BIAS_GELU, bias=bh, aux_out=torch.empty(T, H, device=dev, dtype=bf)), 2.0 * T * D * H),
    ("fc2  fwd  +bias+res", lambda: ops.gemm(xh, wfc2, torch.empty(T, D, device=dev, dtype=bf), epi=EPI_BIAS_RES, bias=b1, aux=x), 2.0 * T * D * H),
    ("fc2  dgrad +dgelu [T,D]x[D,H]", lambda: ops.gemm(x, wfc2, torch.empty(T, H, device=dev, dtype=bf), trans_b=True, epi=EPI_DGELU, aux=xh), 2.0 * T * D * H),
    ("fc1  dgrad [T,H]x[H,D]", lambda: ops.gemm(xh, wfc1, torch.empty(T, D, device=dev, dtype=bf), trans_b=True), 2.0 * T * D * H),
    ("qkv  dgrad [T,3D]x[3D,D]", lambda: ops.gemm(dy3, wqkv, torch.empty(T, D, device=dev, dtype=bf), trans_b=True), 2.0 * T * D * 3 * D),
    ("fc1  wgrad [H,T]x[T,D] f32", lambda: ops.gemm(xh, x, torch.zeros(H, D, device=dev), trans_a=True, trans_b=True, epi=EPI_ACCUM, split_k=_wgrad_split(H, D)), 2.0 * T * D * H),
    ("fc2  wgrad [D,T]x[T,H] f32", lambda: ops.gemm(x, xh, torch.zeros(D, H, device=dev), trans_a=True, trans_b=True, epi=EPI_ACCUM, split_k=_wgrad_split(D, H)), 2.0 * T * D * H),
    ("qkv  wgrad [3D,T]x[T,D] f32", lambda: ops.gemm(dy3, x, torch.zeros(3 * D, D, device=dev), trans_a=True, trans_b=True, epi=EPI_ACCUM, split_k=_wgrad_split(3 * D, D)), 2.0 * T * D * 3 * D),
    ("proj wgrad [D,T]x[T,D] f32", lambda: ops.gemm(x, x, torch.zeros(D, D, device=dev), trans_a=True, trans_b=True, epi=EPI_ACCUM, split_k=_wgrad_split(D, D)), 2.0 * T * D * D),
]
if len(sys.argv) > 1 and sys.argv[1] == "probe":
    # shape vs epilogue: the fc1 shape (N = 4096, K = 1024) with the plain epilogues, the qkv shape (N = 3072) with GELU
    cases = [
        ("fc1 shape, no epilogue", lambda: ops.gemm(x, wfc1, torch.empty(T, H, device=dev, dtype=bf)), 2.0 * T * D * H),
        ("fc1 shape, +bias", lambda: ops.gemm(x, wfc1, torch.empty(T, H, device=dev, dtype=bf), epi=EPI_BIAS, bias=bh), 2.0 * T * D * H),
        ("fc1 shape, +bias+gelu", lambda: ops.gemm(x, wfc1, torch.empty(T, H, device=dev, dtype=bf), epi=EPI_BIAS_GELU, bias=bh, aux_out=torch.empty(T, H, device=dev, dtype=bf)), 2.0 * T * D * H),
        ("qkv shape, +bias", lambda: ops.gemm(x, wqkv, torch.empty(T, 3 * D, device=dev, dtype=bf), epi=EPI_BIAS, bias=b3), 2.0 * T * D * 3 * D),
        ("qkv shape, +bias+gelu", lambda: ops.gemm(x, wqkv, torch.empty(T, 3 * D, device=dev, dtype=bf), epi=EPI_BIAS_GELU, bias=b3, aux_out=torch.empty(T, 3 * D, device=dev, dtype=bf)), 2.0 * T * D * 3 * D),
        ("N=2048 shape, +bias", lambda: ops.gemm(x, wfc1[:2048], torch.empty(T, 2048, device=dev, dtype=bf), epi=EPI_BIAS, bias=bh[:2048]), 2.0 * T * D * 2048),
        ("N=8192 (2x fc1 rows), +bias", lambda: ops.gemm(x, torch.cat([wfc1, wfc1]), torch.empty(T, 8192, device=dev, dtype=bf), epi=EPI_BIAS, bias=torch.cat([bh, bh])), 2.0 * T * D * 8192),
    ]
    sys.argv[1:] = []
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, fn, fl in cases:
    if only and only not in name:
        continue
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:36s} {ms:7.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s")
if (not only and len(cases) > 8) or only == "cublas":
    # the library on the SAME shapes (torch -> cuBLASLt, bf16 bias epilogue where it has one): the calibration point
    import torch.nn.functional as F
    lib = [
        ("cuBLASLt qkv fwd +bias", lambda: F.linear(x, wqkv, b3.to(bf)), 2.0 * T * D * 3 * D),
        ("cuBLASLt proj fwd +bias", lambda: F.linear(x, wp, b1.to(bf)), 2.0 * T * D * D),
        ("cuBLASLt fc1 fwd +bias (no gelu)", lambda: F.linear(x, wfc1, bh.to(bf)), 2.0 * T * D * H),
        ("cuBLASLt fc2 fwd +bias", lambda: F.linear(xh, wfc2, b1.to(bf)), 2.0 * T * D * H),
        ("cuBLASLt fc1 dgrad", lambda: xh @ wfc1, 2.0 * T * D * H),
        ("cuBLASLt fc1 wgrad (bf16 out)", lambda: xh.t() @ x, 2.0 * T * D * H),
        ("cuBLASLt qkv wgrad (bf16 out)", lambda: dy3.t() @ x, 2.0 * T * D * 3 * D),
    ]
    for name, fn, fl in lib:
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name:36s} {ms:7.3f} ms {fl / ms / 1e9:8.1f} TFLOP/s")
if not only and len(cases) > 8:
    a = torch.randn(8192, 8192, device=dev).to(bf)
    for _ in range(2):
        a @ a
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        a @ a
    e1.record()
    torch.cuda.synchronize()
    print(f"{'cuBLAS 8192^3 (reference point)':36s} {e0.elapsed_time(e1) / 5:7.3f} ms {2 * 8192 ** 3 / (e0.elapsed_time(e1) / 5) / 1e9:8.1f} TFLOP/s")
