"""Micro-benchmark of the attention kernels at the 117M shape (N=16200, 16 heads x 64): TFLOP/s per kernel."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from orbit2_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=2)
ap.add_argument("--N", type=int, default=16200)
ap.add_argument("--heads", type=int, default=16)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--hd", type=int, default=64)
ap.add_argument("--fwd-only", action="store_true")
ap.add_argument("--two-pass", action="store_true", help="time the deterministic two-kernel backward instead of the one-pass kernel")
a = ap.parse_args()
hd = a.hd
D = a.heads * hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(a.B * a.N, 3 * D, generator=g, device="cuda").to(torch.bfloat16)
dout = torch.randn(a.B * a.N, D, generator=g, device="cuda").to(torch.bfloat16)
fl = 4.0 * a.N * a.N * D * a.B
for it in range(2):
    out, lse = ops.attn_fwd(qkv, a.B, a.N, a.heads, hd)
    if not a.fwd_only:
        ops.attn_bwd(qkv, out, dout, lse, a.B, a.N, a.heads, hd, two_pass=a.two_pass)
torch.cuda.synchronize()
ops.TIMERS = {}
for it in range(a.iters):
    out, lse = ops.attn_fwd(qkv, a.B, a.N, a.heads, hd)
    if not a.fwd_only:
        ops.attn_bwd(qkv, out, dout, lse, a.B, a.N, a.heads, hd, two_pass=a.two_pass)
torch.cuda.synchronize()
for name, mult in (("attn_fwd", 1.0), ("attn_bwd_dkv", 2.0), ("attn_bwd_dq", 1.5), ("attn_bwd_fused", 2.5), ("attn_bwd_delta", 0.0),
                   ("attn_bwd_dq_finish", 0.0)):
    if name in ops.TIMERS:
        ms = sum(e0.elapsed_time(e1) for e0, e1 in ops.TIMERS[name]) / len(ops.TIMERS[name])
        print(f"{name:16s} {ms:8.3f} ms  {fl * mult / ms / 1e9:8.1f} TFLOP/s")
