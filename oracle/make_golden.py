"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the LIVE, unmodified reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
The fixtures travel to the GPU box (where /root/reference does not exist) and pin both the
oracle restatement (CPU tests) and the CUDA path (GPU tests) to the reference's own numbers.

For every case we instantiate the reference ``Res_Slim_ViT`` (FusedAttn.NONE = explicit softmax,
dropout/drop-path 0), load the seeded state dict from ``reslim_oracle.init_state_dict`` (so that
var_embed/var_query are non-zero), run forward -> clip_replace_constant -> loss -> backward in
float64, and store inputs, weights, prediction, losses and every parameter gradient as float32/64.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import cases, ref_shim, reslim_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def build_reference_model(ref, cfg, sd, dtype):
    m = ref.Res_Slim_ViT(cfg["default_vars"], cfg["init_img_size"], len(cfg["default_vars"]), cfg["out_channels"],
                         history=1, superres_mag=cfg["superres_mag"], cnn_ratio=cfg["cnn_ratio"],
                         patch_size=cfg["patch_size"], drop_path=0.0, drop_rate=0.0, learn_pos_emb=True,
                         embed_dim=cfg["embed_dim"], depth=cfg["depth"], decoder_depth=cfg["decoder_depth"],
                         num_heads=cfg["num_heads"], mlp_ratio=cfg["mlp_ratio"], FusedAttn_option=ref.FusedAttn.NONE)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    m = m.to(dtype)
    # data_config() needs an initialised process group (res_slimvit.py:159); set what it sets.
    m.spatial_resolution = cfg["spatial_resolution"]
    m.img_size = cfg["img_size"]
    m.in_channels = cfg["in_channels"]
    m.out_channels = cfg["out_channels"]
    m.num_patches = cfg["img_size"][0] * cfg["img_size"][1] // cfg["patch_size"] ** 2
    return m


def ref_clip_replace_constant(y, yhat, out_variables):
    # verbatim semantics of examples/intermediate_downscaling.py:267-278 (in-place on the prediction)
    prcp_index = out_variables.index("total_precipitation_24hr")
    for i in range(yhat.shape[1]):
        if i == prcp_index:
            torch.clamp_(yhat[:, prcp_index, :, :], min=0.0)
    for i in range(yhat.shape[1]):
        if out_variables[i] in O.CONSTANTS:
            yhat[:, i] = y[:, i]
    return yhat


def run_case(name: str, B: int, loss_name: str, seed: int = 0, use_lat: bool = False):
    ref = ref_shim.load_reference()
    cfg = cases.get_case(name)
    dt = torch.float64
    sd = {k: v.to(dt) for k, v in O.init_state_dict(cfg, seed).items()}
    x, y = O.synthetic_batch(cfg, B, cfg["in_vars"], cfg["out_vars"], seed)
    x, y = x.to(dt), y.to(dt)
    m = build_reference_model(ref, cfg, sd, dt)
    m.train()
    yhat = m.forward(x, list(cfg["in_vars"]), list(cfg["out_vars"]))
    pred_raw = yhat.detach().clone()
    yhat = ref_clip_replace_constant(y, yhat, list(cfg["out_vars"]))
    lat = np.linspace(90, -90, yhat.shape[2])
    lw = O.lat_weights(lat) if use_lat else None
    fn = ref.functional
    yc = y[:, :, : yhat.shape[2], : yhat.shape[3]]
    if loss_name == "mae":
        loss_vec = fn.mae(yhat, yc, aggregate_only=False, lat_weights=lw)
    else:
        loss_vec = getattr(fn, loss_name)(yhat, yc, var_names=list(cfg["out_vars"]), var_weights=cfg["var_weights"],
                                          aggregate_only=False, lat_weights=lw)
    loss = loss_vec[-1]
    loss.backward()
    out = {"x": x.float().numpy(), "y": y.float().numpy(), "pred_raw": pred_raw.numpy(),
           "pred": yhat.detach().numpy(), "loss_vec": loss_vec.detach().numpy(), "lat": lat,
           "meta": np.array([name, str(B), loss_name, str(seed), str(int(use_lat))])}
    for k, v in m.state_dict().items():
        out["w/" + k] = v.float().numpy()
    for k, p in m.named_parameters():
        out["g/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().astype(np.float32)
    return out


# BASELINE.json configs[0] (interm_8m, 5.5 M parameters): the full fixture would be 40+ MB, so this one keeps the inputs,
# the prediction, the loss vector and the gradients of a parameter subset that touches every part of the path; the weights
# are the seeded reference init (O.init_state_dict) and only their checksum is stored.
COMPACT_GRADS = ["var_query", "var_embed", "spatial_embed.weight", "token_embeds.3.proj.weight", "var_agg.kv.weight",
                 "var_agg.proj.bias", "blocks.0.norm1.weight", "blocks.0.attn.qkv.bias", "blocks.2.attn.proj.weight",
                 "blocks.5.mlp.fc1.bias", "blocks.5.mlp.fc2.weight", "norm.bias", "head.0.weight", "head.8.bias",
                 "path2.0.weight", "path2.3.bias", "conv_out.weight"]


def weight_checksum(sd) -> np.ndarray:
    return np.array([float(v.double().abs().sum()) for _, v in sorted(sd.items())])


def run_case_compact(name: str, B: int, loss_name: str, seed: int, use_lat: bool):
    full = run_case(name, B, loss_name, seed, use_lat)
    out = {k: full[k] for k in ("x", "y", "pred", "loss_vec", "lat", "meta")}
    out["pred"] = out["pred"].astype(np.float32)
    for k in COMPACT_GRADS:                   # matrices: the first 8 rows only ("g8/"), everything else in full ("g/")
        g = full["g/" + k]
        if g.ndim == 2 and g.shape[0] > 8 and g.size > 4096:
            out["g8/" + k] = g[:8].copy()
        else:
            out["g/" + k] = g
    out["w_checksum"] = weight_checksum({k[2:]: torch.from_numpy(v) for k, v in full.items() if k.startswith("w/")})
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    out = run_case_compact("8m", 1, "bayesian_tv", 0, True)
    fn = os.path.join(OUT, "8m_bayesian_tv_lat_compact.npz")
    np.savez_compressed(fn, **out)
    print(fn, os.path.getsize(fn) // 1024, "KiB", "loss", out["loss_vec"])
    jobs = [("tiny", 2, "mse", False), ("tiny", 2, "bayesian_tv", True), ("tiny_prism", 1, "mae", True)]
    for name, B, loss_name, use_lat in jobs:
        out = run_case(name, B, loss_name, 0, use_lat)
        fn = os.path.join(OUT, f"{name}_{loss_name}{'_lat' if use_lat else ''}.npz")
        np.savez_compressed(fn, **out)
        print(fn, os.path.getsize(fn) // 1024, "KiB", "loss", out["loss_vec"])
    # loss-only vectors (no model): reference functional mse / mae / bayesian_tv on a fixed field
    ref = ref_shim.load_reference()
    g = torch.Generator().manual_seed(7)
    pred = torch.randn(2, 3, 24, 40, generator=g, dtype=torch.float64)
    tgt = torch.randn(2, 3, 24, 40, generator=g, dtype=torch.float64)
    lw = O.lat_weights(np.linspace(90, -90, 24))
    names = cases.OUT_VARS_3
    d = {"pred": pred.numpy(), "target": tgt.numpy(), "lat": np.linspace(90, -90, 24)}
    for use_lat in (False, True):
        sfx = "_lat" if use_lat else ""
        w = lw if use_lat else None
        for nm in ("mse", "bayesian_tv"):
            p = pred.clone().requires_grad_(True)
            v = getattr(ref.functional, nm)(p, tgt, var_names=names, var_weights=cases.VAR_WEIGHTS,
                                            aggregate_only=False, lat_weights=w)
            v[-1].backward()
            d[nm + sfx] = v.detach().numpy(); d[nm + sfx + "_grad"] = p.grad.numpy()
        p = pred.clone().requires_grad_(True)
        v = ref.functional.mae(p, tgt, aggregate_only=False, lat_weights=w)
        v[-1].backward()
        d["mae" + sfx] = v.detach().numpy(); d["mae" + sfx + "_grad"] = p.grad.numpy()
    fn = os.path.join(OUT, "loss_vectors.npz")
    np.savez_compressed(fn, **d)
    print(fn, os.path.getsize(fn) // 1024, "KiB")


if __name__ == "__main__":
    main()
