"""A few launches of the fused loss kernels at the interm_117m shapes (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from orbit2_b200 import _lib as L, ops  # noqa: E402

B, C, Ho, Wo = 8, 3, 720, 1440
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
pred = torch.randn(B, C, Ho, Wo, generator=g, device=dev).to(torch.bfloat16)
tgt = torch.randn(B, C, Ho, Wo, generator=g, device=dev)
chw = torch.tensor([1.0, 10.0, 10.0], device=dev)
lat = torch.ones(Ho, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ops.loss_fwd_bwd(pred, tgt, L.LOSS_BAYESIAN_TV, ch_w=chw, lat_w=lat, clamp_ch=0)
    ops.loss_fwd_bwd(pred, tgt, L.LOSS_MSE, ch_w=chw, lat_w=lat, clamp_ch=0)
torch.cuda.synchronize()
