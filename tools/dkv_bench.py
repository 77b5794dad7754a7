"""dkv-only micro-benchmark (A/B of kernel variants through O2B200_LIB)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from orbit2_b200 import ops, _lib as L  # noqa: E402
B, N, heads, hd = 8, 16200, 16, 64
D = heads * hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda").to(torch.bfloat16)
dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
out, lse = ops.attn_fwd(qkv, B, N, heads, hd)
lib = L.load()
dqkv = torch.empty_like(qkv); delta = torch.empty(B, heads, N, device="cuda")
args = (ops._ptr(qkv), ops._ptr(out), ops._ptr(dout), ops._ptr(lse), ops._ptr(dqkv), ops._ptr(delta), B, N, heads, hd, hd ** -0.5, ops._stream())
part = int(os.environ.get("PART", "2"))
lib.o2_attn_bwd_parts(1, 1, *args)
for _ in range(2):
    lib.o2_attn_bwd_parts(1, part, *args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
it = 6
e0.record()
for _ in range(it):
    lib.o2_attn_bwd_parts(1, part, *args)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / it
mult = {2: 2.0, 4: 1.5}[part]
print(f"part {part}: {ms:.3f} ms  {4.0 * N * N * D * B * mult / ms / 1e9:.1f} TFLOP/s")
