"""Shared helpers for the parity tests (oracle side = test infrastructure)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = [str(s) for s in z["meta"]]
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.keys() if k.startswith("w/")}
    grads = {k[2:]: torch.from_numpy(z[k]) for k in z.keys() if k.startswith("g/")}
    return z, meta, sd, grads


def build_model(cfg, sd=None, device="cpu", compute_dtype=None, drop_rate=0.0, drop_path=0.0):
    from orbit2_b200.reslim import Res_Slim_ViT
    m = Res_Slim_ViT(cfg["default_vars"], cfg["init_img_size"], len(cfg["default_vars"]), cfg["out_channels"], 1,
                     superres_mag=cfg["superres_mag"], cnn_ratio=cfg["cnn_ratio"], patch_size=cfg["patch_size"],
                     drop_path=drop_path, drop_rate=drop_rate, learn_pos_emb=True, embed_dim=cfg["embed_dim"],
                     depth=cfg["depth"], decoder_depth=cfg["decoder_depth"], num_heads=cfg["num_heads"],
                     mlp_ratio=cfg["mlp_ratio"], compute_dtype=compute_dtype)
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    m.spatial_resolution = cfg["spatial_resolution"]
    m.img_size = tuple(cfg["img_size"])
    return m.to(device)


def reference_schedule_bf16_grads(cfg, sd, x, y, loss_name, lat_w=None):
    """Parameter gradients of the REFERENCE SCHEDULE itself in bf16 on this GPU: the oracle port under
    torch.autocast(bf16) with fp32 master weights (what intermediate_downscaling.py:601-607 runs).  Its error against the
    float64 oracle is the arithmetic's own floor for each parameter; the bf16 parity tests bound ours by
    max(2e-2, 2 x that floor) wherever a blanket 2e-2 does not hold (tools/bf16_grad_error.py prints the full table)."""
    from oracle import reslim_oracle as O
    sdc = {k: v.float().cuda().requires_grad_(True) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = O.training_step(sdc, cfg, x.float().cuda(), y.float().cuda(), cfg["in_vars"], cfg["out_vars"], loss_name,
                               cfg["var_weights"], lat_w.float().cuda() if lat_w is not None else None)
    loss.float().backward()
    return {k: v.grad.double().cpu() for k, v in sdc.items() if v.grad is not None}
