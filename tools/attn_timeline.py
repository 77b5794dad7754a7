"""Debug helper: builds a -DO2_TIMELINE copy of the library (under /tmp) and prints the clock64() timeline of the dkv
kernel's hand-off points for a few steady-state sub-tiles of CTA (0,0)."""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
src = sorted(glob.glob(os.path.join(ROOT, "orbit2_b200", "csrc", "*.cu")))
out = "/tmp/libo2b200_tl.so"
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-DO2_TIMELINE",
       "-Xcompiler", "-fPIC", "-shared", "-o", out, *src, "-lcudart_static"]
subprocess.check_call(cmd)
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from orbit2_b200 import _lib  # noqa: E402
_lib.LIB_PATH = out
from orbit2_b200 import ops  # noqa: E402

B, N, heads, hd = 1, 16200, 16, 64
D = heads * hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda").to(torch.bfloat16)
dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
which = sys.argv[1] if len(sys.argv) > 1 else "dkv"
for _ in range(2):
    o, lse = ops.attn_fwd(qkv, B, N, heads, hd)
if which != "fwd":
    ops.attn_bwd(qkv, o, dout, lse, B, N, heads, hd)
torch.cuda.synchronize()
lib = _lib.load()
n = 8 * 2 * 16
buf = (C.c_longlong * n)()
lib.o2_debug_timeline.argtypes = [C.c_void_p, C.c_int]
print("rc", lib.o2_debug_timeline(buf, n))
v = list(buf)
t0 = min(x for x in v if x > 0)
which = sys.argv[1] if len(sys.argv) > 1 else "dkv"
names = {0: "mma:wait_pd", 1: "mma:got_pd", 2: "mma:issued", 3: "mma:dvdk_issued", 9: "mma:qdo_present", 10: "prod:want_stage(tile u/2)", 11: "prod:done(tile u/2)", 4: "sm:wait_sd", 5: "sm:got_sd", 6: "sm:ld_done", 7: "sm:math_done", 8: "sm:arrived"}
ev = []
for u in range(8):
    for t in range(2):
        for s, nm in names.items():
            x = v[(u * 2 + t) * 16 + s]
            if x:
                ev.append((x - t0, u, t, nm))
for x, u, t, nm in sorted(ev):
    print(f"{x:8d}  u={u} t={t} {nm}")
