"""One launch of each head-tail / residual-branch kernel at the interm_117m shapes (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from orbit2_b200 import ops  # noqa: E402

B, V, C, H, W, p, mag = 8, 23, 3, 180, 360, 2, 4
gh, gw = H // p, W // p
T = B * gh * gw
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device=dev)
x = rn(B, V, H, W)
idx = [21, 6, 5, 0, 1, 2, 3]
w1, b1 = rn(64, 7, 3, 3) * 0.1, rn(64) * 0.1
w2, b2, wo, bo = rn(C, 4, 3, 3) * 0.1, rn(C) * 0.1, rn(C, C, 3, 3) * 0.1, rn(C) * 0.1
ho = rn(T, C * (mag * p) ** 2).to(bf)
dp = rn(B, C, H * mag, W * mag).to(bf)
G = [torch.zeros_like(t) for t in (wo, bo, w2, b2)]
Gw = [torch.zeros_like(w1), torch.zeros_like(b1)]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    h1, g1 = ops.path2_conv1_fwd(x, idx, w1, b1, bf, mag=mag)
    ops.headtail_fwd(ho, h1, wo, bo, w2, b2, B, C, gh, gw, p, mag, g1=g1)
    _, dh1 = ops.headtail_bwd(dp, ho, h1, wo, w2, *G, B, C, gh, gw, p, mag, g1=g1)
    ops.path2_conv1_bwd(x, idx, dh1, *Gw)
torch.cuda.synchronize()
