"""Head-dim-256 attention (interm_10b: 32 heads x 256, L = 512) forward / backward timing: the fp32 SIMT arm and the bf16
tcgen05 kernels (csrc/attn_tc256.cu).   python tools/attn256_bench.py [B N heads hd]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from orbit2_b200 import ops  # noqa: E402

B, N, heads, hd = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (8, 512, 32, 256)))
D = heads * hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda")
dout = torch.randn(B * N, D, generator=g, device="cuda")


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fl = 4.0 * B * heads * N * N * hd
for name, (x, dy) in (("fp32 SIMT", (qkv, dout)), ("bf16 tcgen05", (qkv.bfloat16(), dout.bfloat16()))):
    out, lse = ops.attn_fwd(x, B, N, heads, hd)
    tf = timed(lambda: ops.attn_fwd(x, B, N, heads, hd))
    tb = timed(lambda: ops.attn_bwd(x, out, dy, lse, B, N, heads, hd))
    print(f"{name:13s} hd={hd} B={B} N={N} heads={heads}: fwd {tf:.3f} ms {fl / tf / 1e9:.1f} TFLOP/s   "
          f"bwd {tb:.3f} ms {2.5 * fl / tb / 1e9:.1f} TFLOP/s (algorithmic)")
