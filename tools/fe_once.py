"""A few launches of the front-end kernels at the interm_117m shapes (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from orbit2_b200 import ops  # noqa: E402

B, V, H, W, p, heads, hd, D = 8, 23, 180, 360, 2, 16, 64, 1024
gh, gw = H // p, W // p
T = B * gh * gw
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device=dev)
x = rn(B, V, H, W)
ts, tv = rn(V, heads, 5), rn(heads, V * 5, hd) * 0.1
do = rn(T, D).to(bf)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ops.frontend_fwd(x, ts, tv, p, gh, gw, hd, bf)
    ops.frontend_bwd(x, ts, tv, do, p, gh, gw, hd)
torch.cuda.synchronize()
