"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the *oracle*) of ORBIT-2's Reslim hot path.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product package
``orbit2_b200`` never imports anything under ``oracle/``.

Parity pinning: the reference's own tests hold no golden vectors for this path (SURVEY.md section 4
/ 8c) -- "parity unpinned by reference tests".  We pin the restatement ourselves:
``tests/test_oracle_vs_reference.py`` runs it against the live, unmodified reference module
(imported through ``oracle/ref_shim.py``) in the build container, and ``oracle/make_golden.py``
writes reference outputs/gradients to ``tests/golden/*.npz`` which travel to the GPU box.

Everything is plain functional PyTorch over a ``state_dict`` with the reference's key names, written
"as the reference computes it" (per-variable conv patch embed, materialised [B,V,L,D] tensor,
explicit softmax), in whatever dtype the inputs have (float64 for ground truth).  Each function cites
the reference lines it restates; paths are relative to ``/root/reference``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# src/climate_learn/data/processing/era5_constants.py:83
CONSTANTS = ["orography", "land_sea_mask", "slt", "lattitude", "longitude"]
STATIC_VARS = ["land_sea_mask", "orography", "lattitude", "landcover"]


# ----------------------------------------------------------------------------------------------
# position embedding: src/climate_learn/models/hub/components/pos_embed.py:20-67
# ----------------------------------------------------------------------------------------------
def sincos_1d(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    omega = np.arange(embed_dim // 2, dtype=float) / (embed_dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim: int, gh: int, gw: int) -> np.ndarray:
    """[gh*gw, D]; first half encodes the *w* coordinate (meshgrid(w, h), 'w goes first')."""
    grid_h = np.arange(gh, dtype=float)
    grid_w = np.arange(gw, dtype=float)
    grid = np.stack(np.meshgrid(grid_w, grid_h), axis=0).reshape(2, 1, gh, gw)
    return np.concatenate([sincos_1d(embed_dim // 2, grid[0]), sincos_1d(embed_dim // 2, grid[1])], axis=1)


def interp_pos_embed(pos_embed: torch.Tensor, patch_size: int, img_size) -> torch.Tensor:
    """pos_embed.py:103-138 -- bicubic resample when the init grid differs (assumes W/H == 2)."""
    D = pos_embed.shape[-1]
    n = pos_embed.shape[-2]
    oh = int((n // 2) ** 0.5)
    ow = 2 * oh
    nh, nw = img_size[0] // patch_size, img_size[1] // patch_size
    if oh == nh:
        return pos_embed
    t = pos_embed.reshape(-1, oh, ow, D).permute(0, 3, 1, 2)
    t = F.interpolate(t, size=(nh, nw), mode="bicubic", align_corners=False)
    return t.permute(0, 2, 3, 1).flatten(1, 2)


# ----------------------------------------------------------------------------------------------
# model forward: src/climate_learn/models/hub/res_slimvit.py
# ----------------------------------------------------------------------------------------------
def find_var_index(in_vars: Sequence[str], out_vars: Sequence[str]) -> List[int]:
    """res_slimvit.py:302-310 (raises ValueError when a static field is missing)."""
    in_vars = list(in_vars)
    return [in_vars.index(v) for v in out_vars] + [in_vars.index(v) for v in STATIC_VARS]


def path2(sd: Dict[str, torch.Tensor], x7: torch.Tensor, mag: int) -> torch.Tensor:
    """res_slimvit.py:107-112: conv3x3 -> GELU(erf) -> PixelShuffle(mag) -> conv3x3."""
    h = F.conv2d(x7, sd["path2.0.weight"], sd["path2.0.bias"], padding=1)
    h = F.gelu(h)
    h = F.pixel_shuffle(h, mag)
    return F.conv2d(h, sd["path2.3.weight"], sd["path2.3.bias"], padding=1)


def var_agg_attention(sd, var_query, x, num_heads):
    """components/attention.py:132-183 with FusedAttn.NONE (explicit softmax), tp=1.
    var_query [T,1,D], x [T,V,D] -> [T,1,D]."""
    T, V, D = x.shape
    hd = D // num_heads
    q = F.linear(var_query, sd["var_agg.q.weight"]).reshape(T, 1, num_heads, hd).permute(0, 2, 1, 3)
    kv = F.linear(x, sd["var_agg.kv.weight"]).reshape(T, V, 2, num_heads, hd).permute(2, 0, 3, 1, 4)
    k, v = kv.unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(T, 1, D)
    return F.linear(o, sd["var_agg.proj.weight"], sd["var_agg.proj.bias"])


# attention.py:54-78 offers three numerically equivalent back ends; parity tests use the explicit softmax
# (FusedAttn.NONE), the CPU timing baseline uses F.scaled_dot_product_attention (FusedAttn.DEFAULT, :66-71) --
# the only back end of the reference that is practical on host cores at L = 16200 (no [L, L] matrix).
USE_SDPA = False


def block(sd, pre: str, x, num_heads, masks=None):
    """components/vit_blocks.py:76-81, attention.py:43-87 (NONE / DEFAULT path), mlp.py:57-73.
    ``masks`` (training mode with dropout): dict of PRE-SCALED keep masks (nn.Dropout: mask / (1 - p); timm DropPath:
    per-sample bernoulli(keep) / keep), keys pre + {attn [B,h,N,N], proj [B,N,D], path1 [B,1,1], drop1 [B,N,4D],
    drop2 [B,N,D], path2 [B,1,1]}; a missing key means the site is inactive."""
    m = masks or {}
    B, N, D = x.shape
    hd = D // num_heads
    h = F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5)
    qkv = F.linear(h, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    if USE_SDPA and (pre + "attn") not in m:
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, D)
    else:
        attn = ((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
        if (pre + "attn") in m:
            attn = attn * m[pre + "attn"]                                   # attention.py:75 attn_drop
        o = (attn @ v).transpose(1, 2).reshape(B, N, D)
    br = F.linear(o, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    if (pre + "proj") in m:
        br = br * m[pre + "proj"]                                           # attention.py:81 proj_drop
    if (pre + "path1") in m:
        br = br * m[pre + "path1"]                                          # vit_blocks.py:78 drop_path1
    x = x + br
    h = F.layer_norm(x, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
    h = F.gelu(F.linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))
    if (pre + "drop1") in m:
        h = h * m[pre + "drop1"]                                            # mlp.py:65 drop1
    br = F.linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    if (pre + "drop2") in m:
        br = br * m[pre + "drop2"]                                          # mlp.py:68 drop2
    if (pre + "path2") in m:
        br = br * m[pre + "path2"]                                          # vit_blocks.py:79 drop_path2
    return x + br


def unpatchify(x: torch.Tensor, img_size, patch_size: int, scaling: int, c: int) -> torch.Tensor:
    """res_slimvit.py:167-179 -- a flat row-major reinterpretation (p = patch_size, NOT mag*p)."""
    p = patch_size
    h = img_size[0] * scaling // p
    w = img_size[1] * scaling // p
    x = x.reshape(x.shape[0], h, w, p, p, c)
    x = torch.einsum("nhwpqc->nchpwq", x)
    return x.reshape(x.shape[0], c, h * p, w * p)


def forward(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, in_vars: Sequence[str],
            out_vars: Sequence[str], taps: Optional[dict] = None, masks: Optional[dict] = None) -> torch.Tensor:
    """Res_Slim_ViT.forward, res_slimvit.py:312-338.  ``masks``: pre-scaled dropout / drop-path keep masks of a
    training-mode forward (see ``block``; "pos" [B,L,D] is pos_drop, res_slimvit.py:284); None = eval / p = 0.

    cfg keys: default_vars, img_size (current input grid), patch_size, superres_mag, num_heads, depth,
    decoder_depth, spatial_resolution.
    """
    if x.dim() == 5:
        x = x.flatten(1, 2)
    default_vars = list(cfg["default_vars"])
    p = cfg["patch_size"]
    mag = cfg["superres_mag"]
    heads = cfg["num_heads"]
    C = len(out_vars)
    B = x.shape[0]

    idx = find_var_index(in_vars, out_vars)
    p2 = path2(sd, x[:, idx], mag)                                      # :233-242

    # forward_encoder :245-299
    var_ids = [default_vars.index(v) for v in in_vars]                   # var_map, :182-201 (KeyError -> ValueError)
    embeds = []
    for i, vid in enumerate(var_ids):                                    # :254-257, patch_embed.py:47-53
        e = F.conv2d(x[:, i:i + 1], sd[f"token_embeds.{vid}.proj.weight"], sd[f"token_embeds.{vid}.proj.bias"], stride=p)
        embeds.append(e.flatten(2).transpose(1, 2))
    t = torch.stack(embeds, dim=1)                                       # B,V,L,D
    t = t + sd["var_embed"][:, var_ids, :].unsqueeze(2)                  # :260-262
    b, V, L, D = t.shape
    t = torch.einsum("bvld->blvd", t).flatten(0, 1)                      # :211-212
    vq = sd["var_query"].expand(t.shape[0], -1, -1)
    t = var_agg_attention(sd, vq, t, heads).squeeze(1).unflatten(0, (b, L))   # :216-228
    if taps is not None:
        taps["agg"] = t
    t = t + interp_pos_embed(sd["pos_embed"], p, cfg["img_size"])       # :270-273
    res = torch.tensor([float(cfg["spatial_resolution"])], dtype=t.dtype, device=t.device)
    t = t + F.linear(res, sd["spatial_embed.weight"], sd["spatial_embed.bias"])[None, None]  # :277-281
    if taps is not None:
        taps["tokens0"] = t
    if masks is not None and "pos" in masks:
        t = t * masks["pos"]                                             # :284 pos_drop
    for i in range(cfg["depth"]):                                        # :291-292
        t = block(sd, f"blocks.{i}.", t, heads, masks)
        if taps is not None:
            taps[f"block{i}"] = t
    t = F.layer_norm(t, (D,), sd["norm.weight"], sd["norm.bias"], 1e-5)  # :294

    for j in range(cfg["decoder_depth"]):                                # head :115-120
        t = F.gelu(F.linear(t, sd[f"head.{2 * j}.weight"], sd[f"head.{2 * j}.bias"]))
    j = 2 * cfg["decoder_depth"]
    t = F.linear(t, sd[f"head.{j}.weight"], sd[f"head.{j}.bias"])
    if taps is not None:
        taps["head"] = t
    img = unpatchify(t, cfg["img_size"], p, mag, C)                      # :329
    img = F.conv2d(img, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)   # :331
    return img + p2[:, :, : img.shape[2], : img.shape[3]]                # :333-336


# ----------------------------------------------------------------------------------------------
# driver glue: examples/intermediate_downscaling.py
# ----------------------------------------------------------------------------------------------
def clip_replace_constant(y: torch.Tensor, yhat: torch.Tensor, out_vars: Sequence[str]) -> torch.Tensor:
    """intermediate_downscaling.py:267-278 (out-of-place restatement so autograd can see through it;
    the reference clamps in place on a view, which has the same gradient: 0 where clamped).
    Raises ValueError if 'total_precipitation_24hr' is not an output variable, like the reference."""
    out_vars = list(out_vars)
    pi = out_vars.index("total_precipitation_24hr")
    chans = []
    for i in range(yhat.shape[1]):
        c = yhat[:, i]
        if i == pi:
            c = torch.clamp(c, min=0.0)
        if out_vars[i] in CONSTANTS:
            c = y[:, i].to(c.dtype)
        chans.append(c)
    return torch.stack(chans, dim=1)


def lat_weights(lat: np.ndarray) -> torch.Tensor:
    """metrics/metrics.py:58-65 -> float64 [1,1,H,1]."""
    w = np.cos(np.deg2rad(np.asarray(lat, dtype=np.float64)))
    w = w / w.mean()
    return torch.from_numpy(w).view(1, 1, -1, 1)


def _weight_and_reduce(error, pred, var_names, var_weights, aggregate_only, lat_w):
    if lat_w is not None:
        error = error * lat_w
    if var_names is not None:
        assert len(var_names) == pred.shape[1]
        cw = torch.ones(pred.shape[1], device=pred.device, dtype=pred.dtype)
        for i, v in enumerate(var_names):
            cw[i] = var_weights.get(v, 1.0)
        error = error * cw.view(1, -1, 1, 1)
    per_ch = error.mean([0, 2, 3])
    loss = error.mean()
    return loss if aggregate_only else torch.cat((per_ch, loss.unsqueeze(0)))


def mse(pred, target, var_names=None, var_weights=None, aggregate_only=False, lat_w=None):
    """metrics/functional.py:173-202."""
    return _weight_and_reduce((pred - target).square(), pred, var_names, var_weights, aggregate_only, lat_w)


def mae(pred, target, aggregate_only=False, lat_w=None):
    """metrics/functional.py:218-232 (no variable weights)."""
    return _weight_and_reduce((pred - target).abs(), pred, None, None, aggregate_only, lat_w)


def bayesian_tv(pred, target, var_names=None, var_weights=None, aggregate_only=False, lat_w=None):
    """metrics/functional.py:117-167: MSE + 0.02*(|dv| + |dh| + 0.7|d_diag| + 0.7|d_anti|) of pred."""
    e = (pred - target).square()
    d1 = F.pad((pred[:, :, 1:, :] - pred[:, :, :-1, :]).abs(), (0, 0, 0, 1))
    d2 = F.pad((pred[:, :, :, 1:] - pred[:, :, :, :-1]).abs(), (0, 1))
    d3 = F.pad((pred[:, :, 1:, 1:] - pred[:, :, :-1, :-1]).abs(), (0, 1, 0, 1))
    d4 = F.pad((pred[:, :, 1:, :-1] - pred[:, :, :-1, 1:]).abs(), (1, 0, 0, 1))
    e = e + 0.02 * (d1 + d2 + 0.7 * d3 + 0.7 * d4)
    return _weight_and_reduce(e, pred, var_names, var_weights, aggregate_only, lat_w)


LOSSES = {"mse": mse, "bayesian_tv": bayesian_tv}


def training_step(sd, cfg, x, y, in_vars, out_vars, loss_name="mse", var_weights=None, lat_w=None,
                  taps: Optional[dict] = None, masks: Optional[dict] = None) -> torch.Tensor:
    """intermediate_downscaling.py:281-306: forward, clip/replace, crop target, loss (aggregate)."""
    yhat = forward(sd, cfg, x, in_vars, out_vars, taps, masks)
    yhat = clip_replace_constant(y, yhat, out_vars)
    if taps is not None:
        taps["preds"] = yhat
    yc = y[:, :, : yhat.shape[2], : yhat.shape[3]]
    if loss_name == "mae":
        return mae(yhat, yc, True, lat_w)
    return LOSSES[loss_name](yhat, yc, list(out_vars), var_weights or {}, True, lat_w)


# ----------------------------------------------------------------------------------------------
# weights and synthetic data (shared by tests / bench so both sides see identical inputs)
# ----------------------------------------------------------------------------------------------
def param_shapes(cfg: dict) -> Dict[str, tuple]:
    """State-dict ABI of Res_Slim_ViT, in the reference's registration order."""
    D = cfg["embed_dim"]; p = cfg["patch_size"]; Vd = len(cfg["default_vars"]); C = cfg["out_channels"]
    mag = cfg["superres_mag"]; cr = cfg.get("cnn_ratio", 4); Hd = int(D * cfg.get("mlp_ratio", 4.0))
    L0 = (cfg["init_img_size"][0] // p) * (cfg["init_img_size"][1] // p)
    s = {"var_embed": (1, Vd, D), "var_query": (1, 1, D), "pos_embed": (1, L0, D),
         "spatial_embed.weight": (D, 1), "spatial_embed.bias": (D,)}
    for i in range(Vd):
        s[f"token_embeds.{i}.proj.weight"] = (D, 1, p, p)
        s[f"token_embeds.{i}.proj.bias"] = (D,)
    s.update({"var_agg.q.weight": (D, D), "var_agg.kv.weight": (2 * D, D), "var_agg.proj.weight": (D, D),
              "var_agg.proj.bias": (D,)})
    for i in range(cfg["depth"]):
        b = f"blocks.{i}."
        s.update({b + "norm1.weight": (D,), b + "norm1.bias": (D,), b + "attn.qkv.weight": (3 * D, D),
                  b + "attn.qkv.bias": (3 * D,), b + "attn.proj.weight": (D, D), b + "attn.proj.bias": (D,),
                  b + "norm2.weight": (D,), b + "norm2.bias": (D,), b + "mlp.fc1.weight": (Hd, D),
                  b + "mlp.fc1.bias": (Hd,), b + "mlp.fc2.weight": (D, Hd), b + "mlp.fc2.bias": (D,)})
    s.update({"norm.weight": (D,), "norm.bias": (D,), "path2.0.weight": (cr * mag * mag, C + 4, 3, 3),
              "path2.0.bias": (cr * mag * mag,), "path2.3.weight": (C, cr, 3, 3), "path2.3.bias": (C,)})
    for j in range(cfg["decoder_depth"]):
        s[f"head.{2 * j}.weight"] = (D, D); s[f"head.{2 * j}.bias"] = (D,)
    j = 2 * cfg["decoder_depth"]
    s[f"head.{j}.weight"] = (C * (mag * p) ** 2, D); s[f"head.{j}.bias"] = (C * (mag * p) ** 2,)
    s["conv_out.weight"] = (C, C, 3, 3); s["conv_out.bias"] = (C,)
    return s


def init_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Reference-style init (res_slimvit.py:125-145): Linear trunc_normal(.02)/bias 0, LN 1/0, convs
    kaiming-uniform(a=sqrt5), pos_embed sin-cos; var_embed/var_query ~ N(0,.02) instead of the
    reference's zeros so that front-end bugs are visible.  Not RNG-identical to
    the reference constructor -- parity tests load the *same* tensors on both sides."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if k == "pos_embed":
            p = cfg["patch_size"]
            t = torch.from_numpy(sincos_2d(shp[-1], cfg["init_img_size"][0] // p, cfg["init_img_size"][1] // p)).float()[None]
        elif k in ("var_embed", "var_query"):
            t = torch.randn(shp, generator=g) * 0.02
        elif k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "norm.weight":
            t = torch.ones(shp) + 0.1 * torch.randn(shp, generator=g)      # perturbed so LN affine is exercised
        elif k.endswith("norm1.bias") or k.endswith("norm2.bias") or k == "norm.bias":
            t = 0.05 * torch.randn(shp, generator=g)
        elif len(shp) == 4:                                               # conv weight
            fan_in = shp[1] * shp[2] * shp[3]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif len(shp) == 2:                                               # linear weight
            t = torch.nn.init.trunc_normal_(torch.empty(shp), std=0.02, generator=g)
        else:                                                             # biases: small non-zero
            t = 0.02 * torch.randn(shp, generator=g)
        sd[k] = t.to(dtype)
    return sd


def synthetic_batch(cfg: dict, B: int, in_vars: Sequence[str], out_vars: Sequence[str], seed: int = 0,
                    hi_rows: Optional[int] = None):
    """SURVEY.md section 8d synthetic inputs: randn fields, precip >= 0 (log1p like LogTransform),
    land_sea_mask in {0,1}; target randn with precip >= 0."""
    g = torch.Generator().manual_seed(seed + 1000)
    H, W = cfg["img_size"]
    mag = cfg["superres_mag"]
    x = torch.randn(B, len(in_vars), H, W, generator=g)
    in_vars = list(in_vars)
    if "total_precipitation_24hr" in in_vars:
        i = in_vars.index("total_precipitation_24hr")
        x[:, i] = torch.log1p(torch.relu(x[:, i]) * 2.0)
    if "land_sea_mask" in in_vars:
        i = in_vars.index("land_sea_mask")
        x[:, i] = (x[:, i] > 0).float()
    Ho = hi_rows if hi_rows is not None else H * mag
    y = torch.randn(B, len(out_vars), Ho, W * mag, generator=g)
    out_vars = list(out_vars)
    if "total_precipitation_24hr" in out_vars:
        i = out_vars.index("total_precipitation_24hr")
        y[:, i] = torch.log1p(torch.relu(y[:, i]) * 2.0)
    return x, y


# ----------------------------------------------------------------------------------------------
# either side of the hot path: evaluation metrics and the data pipeline's per-variable transforms
# ----------------------------------------------------------------------------------------------
def rmse(pred, target, aggregate_only=False, lat_w=None):
    """metrics/functional.py:236-257 (no mask)."""
    error = (pred - target).square()
    if lat_w is not None:
        error = error * lat_w
    per_ch = error.mean([2, 3]).sqrt().mean(0)
    loss = per_ch.mean()
    return loss if aggregate_only else torch.cat((per_ch, loss.unsqueeze(0)))


def pearson(pred, target, aggregate_only=False):
    """metrics/functional.py:294-309, :327-337."""
    p = pred.transpose(0, 1).flatten(1)
    t = target.transpose(0, 1).flatten(1)
    p = p - p.mean(1, keepdim=True)
    t = t - t.mean(1, keepdim=True)
    per_ch = F.cosine_similarity(p, t)
    c = per_ch.mean()
    return c if aggregate_only else torch.cat((per_ch, c.unsqueeze(0)))


def mean_bias(pred, target, aggregate_only=False):
    """metrics/functional.py:312-324."""
    per_ch = torch.stack([target[:, i].mean() - pred[:, i].mean() for i in range(pred.shape[1])])
    r = per_ch.mean()
    return r if aggregate_only else torch.cat((per_ch, r.unsqueeze(0)))


def log_transform(x: torch.Tensor) -> torch.Tensor:
    """data/precipmodule.py:21-42 with the pipeline's arguments (m2mm=True, LOG1P=True, thres 0.25 mm/day,
    itermodule.py:208); out of place."""
    t = x * 1000.0
    t = torch.where(t <= 0.25, torch.zeros((), dtype=t.dtype), t)
    return torch.log1p(t)


def normalize_sample(x: torch.Tensor, variables, mean: dict, std: dict, precip=("total_precipitation_24hr", "total_precipitation")):
    """itermodule.py:202-211 + iterdataset.py:360-379: per-variable Normalize(mean, std) or LogTransform; x [.., V, H, W]."""
    out = []
    for i, v in enumerate(variables):
        c = x[..., i, :, :]
        out.append(log_transform(c) if v in precip else (c - float(mean[v][0])) / float(std[v][0]))
    return torch.stack(out, dim=-3)
