"""Drop-in ``Res_Slim_ViT`` for ORBIT-2 whose forward *and* backward run on the libo2b200 kernels.

Mirrors the reference module (src/climate_learn/models/hub/res_slimvit.py:20-338): same constructor
signature (:22-43), attributes read by the driver (:46-60, utils/visualize.py:45-58), ``data_config``
(:148-164), ``forward(x, in_variables, out_variables)`` (:312-338) and the same state-dict keys, so
reference checkpoints load unchanged.  There is no PyTorch/CPU implementation behind it: without the
CUDA library and an sm_100a device every call raises.

Execution plan of one forward (T = B*L tokens, act = fp32 | bf16):
  host prep (tiny, autograd-visible torch ops on parameters only):
      tab_s/tab_v  = exact collapse of patch-embed + var-embed + var-agg q/kv
      posres [L,D] = pos_embed (bicubic-resampled if the grid changed) + spatial_embed(resolution)
  kernels (one manual autograd node, ``ReslimFunction``):
      h1,g1 = o2_path2_conv1_fwd(x[:, idx7])                       residual branch: low-res pre-GELU + activated/shuffled copy
      o     = o2_frontend_fwd(x, tab_s, tab_v)                     [T, D]
      tok   = o2_gemm(o, var_agg.proj) + bias + posres             BIAS_RES epilogue
      per block: LN -> qkv GEMM(+bias) -> flash attention -> proj GEMM(+bias+residual)
                 LN -> fc1 GEMM(+bias, GELU, keeps pre-activation) -> fc2 GEMM(+bias+residual)
      final LN -> head GEMMs (+bias+GELU) -> last head GEMM(+bias)
      preds = o2_headtail_fwd(...)  unpatchify + conv_out + GELU/PixelShuffle/conv2 of h1 + crop-add
"""
from __future__ import annotations

import math
import warnings
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import EPI_ACCUM, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_DGELU, EPI_NONE

STATIC_VARS = ["land_sea_mask", "orography", "lattitude", "landcover"]      # res_slimvit.py:305-308


class FusedAttn:
    """climate_learn/utils/fused_attn.py enum values; accepted and ignored (one attention kernel here)."""
    CK, DEFAULT, NONE = 0, 1, 2


# ------------------------------------------------------------------------------------------------
# parameter containers with the reference's module/attribute names (state-dict ABI)
# ------------------------------------------------------------------------------------------------
class _PatchEmbed(nn.Module):                       # components/patch_embed.py:22-53 (parameters only)
    def __init__(self, img_size, patch_size, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(1, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.num_patches = (img_size[0] // patch_size) * (img_size[1] // patch_size)


class _VarAgg(nn.Module):                           # components/attention.py:92-130
    def __init__(self, dim):
        super().__init__()
        self.q = nn.Linear(dim, dim, bias=False)
        self.kv = nn.Linear(dim, 2 * dim, bias=False)
        self.proj = nn.Linear(dim, dim)


class _Attn(nn.Module):                             # components/attention.py:14-41
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):                              # components/mlp.py:22-55
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class Block(nn.Module):                             # components/vit_blocks.py:34-81 (parameter container)
    def __init__(self, dim, hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _Attn(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, hidden)


def get_2d_sincos_pos_embed(embed_dim, gh, gw):
    """components/pos_embed.py:20-67 ('w goes first')."""
    def one(d, pos):
        omega = 1.0 / 10000 ** (np.arange(d // 2, dtype=float) / (d / 2.0))
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)
    grid = np.stack(np.meshgrid(np.arange(gw, dtype=float), np.arange(gh, dtype=float)), axis=0).reshape(2, 1, gh, gw)
    return np.concatenate([one(embed_dim // 2, grid[0]), one(embed_dim // 2, grid[1])], axis=1)


# ------------------------------------------------------------------------------------------------
# the kernel schedule (shared by the autograd node and the no-autograd training engine)
# ------------------------------------------------------------------------------------------------
class Geometry:
    __slots__ = ("B", "V", "Hx", "Wx", "p", "gh", "gw", "L", "T", "D", "heads", "hd", "depth", "dec", "mag", "C", "cr",
                 "hidden", "idx7", "act", "drop", "ckpt")

    def __repr__(self):
        return "Geometry(" + ", ".join(f"{k}={getattr(self, k)}" for k in self.__slots__ if hasattr(self, k)) + ")"


# dropout site numbering (the keep mask is a hash of (seed, site, element), csrc/dropout.cu)
SITE_POS, SITE_ATTN, SITE_PROJ, SITE_DROP1, SITE_DROP2 = 0, 1, 2, 3, 4


def drop_site(block: int, which: int) -> int:
    return 16 * block + which


class DropPlan:
    """One training-mode forward's dropout / stochastic-depth decisions: the reference's nn.Dropout(p=drop_rate) at
    res_slimvit.py:284, attention.py:81, mlp.py:65,68 and DropPath(dpr[i]) at vit_blocks.py:78-79 with
    dpr = linspace(0, drop_path, depth) (res_slimvit.py:84).  Masks are never stored: the element masks are regenerated
    from ``seed``; the per-sample drop-path factors (bernoulli(keep) / keep, timm semantics) are tiny [B] tensors."""

    def __init__(self, rate: float, dpr: Sequence[float], B: int, seed: int, device):
        self.rate, self.seed = float(rate), int(seed)
        self.dpr, self.B = [float(d) for d in dpr], int(B)
        self.path = []
        for i, sc in enumerate(self._draw(self.seed)):
            self.path.append((None, None) if sc is None else (sc[0].to(device), sc[1].to(device)))

    def _draw(self, seed: int):
        out = []
        for i, dp in enumerate(self.dpr):
            if dp > 0.0:
                gen = torch.Generator().manual_seed((seed + 7919 * (i + 1)) & 0x7FFFFFFFFFFFFFFF)
                keep = 1.0 - dp
                out.append((torch.rand(2, self.B, generator=gen) < keep).to(torch.float32) / keep)
            else:
                out.append(None)
        return out

    def redraw_paths(self, seed: int):
        """CUDA-graph mode: new per-sample drop-path factors written INTO the existing device tensors (the captured
        kernels keep reading the same addresses); the element masks change through the device step word instead
        (o2_dropout_seed_source), ``self.seed`` stays what the graph captured."""
        for (a, b), sc in zip(self.path, self._draw(int(seed))):
            if sc is not None:
                a.copy_(sc[0], non_blocking=True)
                b.copy_(sc[1], non_blocking=True)

    def branch_active(self, i: int) -> bool:
        return self.rate > 0.0 or self.path[i][0] is not None


def kernel_param_names(depth: int, dec: int) -> List[str]:
    """Parameters consumed directly by kernels, in the order ReslimFunction receives them."""
    n = ["var_agg.proj.weight", "var_agg.proj.bias"]
    for i in range(depth):
        b = f"blocks.{i}."
        n += [b + "norm1.weight", b + "norm1.bias", b + "attn.qkv.weight", b + "attn.qkv.bias", b + "attn.proj.weight",
              b + "attn.proj.bias", b + "norm2.weight", b + "norm2.bias", b + "mlp.fc1.weight", b + "mlp.fc1.bias",
              b + "mlp.fc2.weight", b + "mlp.fc2.bias"]
    n += ["norm.weight", "norm.bias"]
    for j in range(dec + 1):
        n += [f"head.{2 * j}.weight", f"head.{2 * j}.bias"]
    n += ["path2.0.weight", "path2.0.bias", "path2.3.weight", "path2.3.bias", "conv_out.weight", "conv_out.bias"]
    return n


def _f32(t):
    return t if t.dtype == torch.float32 else t.float()


def _block_forward(g: Geometry, P, Wc, i: int, tok):
    """One pre-norm block (vit_blocks.py:76-81): returns (output tokens, tensors the block's backward needs).  Dropout /
    drop-path masks are functions of (seed, site, element), so a re-run for activation checkpointing reproduces them."""
    T, D, act = g.T, g.D, g.act
    dev = tok.device
    dp = getattr(g, "drop", None)

    def gemm(a, w, n_out, **kw):
        out = torch.empty(a.shape[0], n_out, device=dev, dtype=act)
        return ops.gemm(a, w, out, **kw)

    b = f"blocks.{i}."
    s = {"x": tok}
    y1, s["mean1"], s["rstd1"] = ops.layernorm_fwd(tok, P[b + "norm1.weight"], P[b + "norm1.bias"])
    qkv = gemm(y1, Wc[b + "attn.qkv.weight"], 3 * D, epi=EPI_BIAS, bias=P[b + "attn.qkv.bias"])
    adrop = (dp.rate, dp.seed, drop_site(i, SITE_ATTN)) if (dp is not None and dp.rate > 0) else None
    ao, s["lse"] = ops.attn_fwd(qkv, g.B, g.L, g.heads, g.hd, adrop)                 # attention.py:75 attn_drop
    fused = act == torch.bfloat16          # the tcgen05 GEMM applies the masks in its epilogue; the fp32 arm in a separate pass
    if dp is not None and dp.branch_active(i) and fused:
        # x + drop_path1(proj_drop(proj(.))): attention.py:81, vit_blocks.py:78
        xm = gemm(ao, Wc[b + "attn.proj.weight"], D, epi=EPI_BIAS_RES, bias=P[b + "attn.proj.bias"], aux=tok,
                  drop=(dp.rate, dp.seed, drop_site(i, SITE_PROJ), dp.path[i][0], g.L))
    elif dp is not None and dp.branch_active(i):
        br = gemm(ao, Wc[b + "attn.proj.weight"], D, epi=EPI_BIAS, bias=P[b + "attn.proj.bias"])
        xm = ops.dropout(br, dp.rate, dp.seed, drop_site(i, SITE_PROJ), res=tok, sample_scale=dp.path[i][0],
                         rows_per_sample=g.L, out=br)
    else:
        xm = gemm(ao, Wc[b + "attn.proj.weight"], D, epi=EPI_BIAS_RES, bias=P[b + "attn.proj.bias"], aux=tok)
    y2, s["mean2"], s["rstd2"] = ops.layernorm_fwd(xm, P[b + "norm2.weight"], P[b + "norm2.bias"])
    pre = torch.empty(T, g.hidden, device=dev, dtype=act)
    drop1 = (dp.rate, dp.seed, drop_site(i, SITE_DROP1), None, 0) if (dp is not None and dp.rate > 0) else None
    h = gemm(y2, Wc[b + "mlp.fc1.weight"], g.hidden, epi=EPI_BIAS_GELU, bias=P[b + "mlp.fc1.bias"], aux_out=pre,
             drop=drop1 if fused else None)                                     # mlp.py:65 drop1
    if drop1 is not None and not fused:
        ops.dropout(h, dp.rate, dp.seed, drop_site(i, SITE_DROP1), out=h)
    if dp is not None and dp.branch_active(i) and fused:
        out = gemm(h, Wc[b + "mlp.fc2.weight"], D, epi=EPI_BIAS_RES, bias=P[b + "mlp.fc2.bias"], aux=xm,
                   drop=(dp.rate, dp.seed, drop_site(i, SITE_DROP2), dp.path[i][1], g.L))   # mlp.py:68 drop2, vit_blocks.py:79
    elif dp is not None and dp.branch_active(i):
        br = gemm(h, Wc[b + "mlp.fc2.weight"], D, epi=EPI_BIAS, bias=P[b + "mlp.fc2.bias"])
        out = ops.dropout(br, dp.rate, dp.seed, drop_site(i, SITE_DROP2), res=xm, sample_scale=dp.path[i][1],
                          rows_per_sample=g.L, out=br)                          # mlp.py:68 drop2, vit_blocks.py:79
    else:
        out = gemm(h, Wc[b + "mlp.fc2.weight"], D, epi=EPI_BIAS_RES, bias=P[b + "mlp.fc2.bias"], aux=xm)
    s.update(y1=y1, qkv=qkv, ao=ao, xm=xm, y2=y2, pre=pre, h=h)
    return out, s


def reslim_forward(g: Geometry, P: Dict[str, torch.Tensor], Wc: Dict[str, torch.Tensor], x, tab_s, tab_v, posres):
    """P: fp32 parameters (biases, LN affine, conv weights); Wc: GEMM weights in the activation dtype.
    Returns (preds, saved) where ``saved`` holds what reslim_backward needs."""
    S = {}
    T, D, act = g.T, g.D, g.act
    dev = x.device

    def gemm(a, w, n_out, **kw):
        out = torch.empty(a.shape[0], n_out, device=dev, dtype=act)
        return ops.gemm(a, w, out, **kw)

    S["h1"], S["g1"] = h1, g1 = ops.path2_conv1_fwd(x, g.idx7, P["path2.0.weight"], P["path2.0.bias"], act, mag=g.mag)
    S["o"] = o = ops.frontend_fwd(x, tab_s, tab_v, g.p, g.gh, g.gw, g.hd, act)
    dp = getattr(g, "drop", None)
    pos_drop = dp is not None and dp.rate > 0                                       # pos_drop, res_slimvit.py:284
    fused = act == torch.bfloat16
    tok = gemm(o, Wc["var_agg.proj.weight"], D, epi=EPI_BIAS_RES, bias=P["var_agg.proj.bias"], aux=posres, aux_rows=g.L,
               drop=(dp.rate, dp.seed, SITE_POS, None, 0, True) if (pos_drop and fused) else None)
    if pos_drop and not fused:
        ops.dropout(tok, dp.rate, dp.seed, SITE_POS, out=tok)
    blocks = []
    ckpt = bool(getattr(g, "ckpt", False))
    for i in range(g.depth):
        tok_in = tok
        tok, s = _block_forward(g, P, Wc, i, tok_in)
        # activation checkpointing (the reference wraps every Block, intermediate_downscaling.py:583-590, 635-637):
        # keep only the block input, reslim_backward re-runs the block's forward kernels to rebuild the rest
        blocks.append({"x": tok_in} if ckpt else s)
    S["blocks"] = blocks
    S["xf"] = tok
    z, S["meanf"], S["rstdf"] = ops.layernorm_fwd(tok, P["norm.weight"], P["norm.bias"])
    zs, pres = [z], []
    for j in range(g.dec):
        pre = torch.empty(T, D, device=dev, dtype=act)
        z = gemm(z, Wc[f"head.{2 * j}.weight"], D, epi=EPI_BIAS_GELU, bias=P[f"head.{2 * j}.bias"], aux_out=pre)
        zs.append(z)
        pres.append(pre)
    S["zs"], S["pres"] = zs, pres
    n_out = g.C * (g.mag * g.p) ** 2
    S["ho"] = ho = gemm(z, Wc[f"head.{2 * g.dec}.weight"], n_out, epi=EPI_BIAS, bias=P[f"head.{2 * g.dec}.bias"])
    preds = ops.headtail_fwd(ho, h1, P["conv_out.weight"], P["conv_out.bias"], P["path2.3.weight"], P["path2.3.bias"],
                             g.B, g.C, g.gh, g.gw, g.p, g.mag, g1=g1)
    return preds, S


def reslim_backward(g: Geometry, P, Wc, x, tab_s, tab_v, S, dpreds, G: Dict[str, torch.Tensor], on_ready=None):
    """ACCUMULATES parameter gradients into the fp32 tensors of ``G`` (the caller zero-fills them once per step) and
    returns (dtab_s, dtab_v, dposres).  ``on_ready(names)`` is
    called as soon as a group of gradients is final (the data-parallel engine starts its all-reduce there)."""
    T, D, act = g.T, g.D, g.act
    dev = x.device

    def dgrad(dy, w, n_in, **kw):
        """dX[T, n_in] = dY[T, n_out] @ W[n_out, n_in]"""
        out = torch.empty(dy.shape[0], n_in, device=dev, dtype=act)
        return ops.gemm(dy, w, out, trans_b=True, **kw)

    def wgrad(dy, a, name):
        """dW[n_out, n_in] += dY^T A  (fp32, split-K over the token dimension so that small weight matrices still
        fill the machine; partial tiles are accumulated with atomics); db += colsum(dY) rides on the same GEMM: its
        epilogue warps add up the dY tiles that stream through shared memory (no separate pass over dY)"""
        w = G[name + ".weight"]
        ops.gemm(dy, a, w, trans_a=True, trans_b=True, epi=EPI_ACCUM, split_k=_wgrad_split(w.shape[0], w.shape[1]),
                 bias=G[name + ".bias"])

    def ready(names):
        if on_ready is not None:
            on_ready(names)

    dho, dh1 = ops.headtail_bwd(dpreds, S["ho"], S["h1"], P["conv_out.weight"], P["path2.3.weight"], G["conv_out.weight"],
                                G["conv_out.bias"], G["path2.3.weight"], G["path2.3.bias"], g.B, g.C, g.gh, g.gw, g.p,
                                g.mag, g1=S["g1"])
    ops.path2_conv1_bwd(x, g.idx7, dh1, G["path2.0.weight"], G["path2.0.bias"])
    ready(["conv_out.weight", "conv_out.bias", "path2.3.weight", "path2.3.bias", "path2.0.weight", "path2.0.bias"])
    del dh1
    zs, pres = S["zs"], S["pres"]
    j = g.dec
    wgrad(dho, zs[j], f"head.{2 * j}")
    if g.dec > 0:
        d = dgrad(dho, Wc[f"head.{2 * j}.weight"], D, epi=EPI_DGELU, aux=pres[j - 1])
    else:
        d = dgrad(dho, Wc[f"head.{2 * j}.weight"], D)
    for j in range(g.dec - 1, -1, -1):
        wgrad(d, zs[j], f"head.{2 * j}")
        if j > 0:
            d = dgrad(d, Wc[f"head.{2 * j}.weight"], D, epi=EPI_DGELU, aux=pres[j - 1])
        else:
            d = dgrad(d, Wc[f"head.{2 * j}.weight"], D)
    ready([f"head.{2 * j}.{s}" for j in range(g.dec + 1) for s in ("weight", "bias")])
    dp = getattr(g, "drop", None)

    def branch_mask(i, which, path_ix):
        """(p, seed, site, sample_scale, rows) of the gradient mask of Block i's residual branch, or None: the LayerNorm
        backward that produces the stream gradient also writes its masked copy (same mask, same scale as the forward)"""
        if i < 0 or dp is None or not dp.branch_active(i):
            return None
        return (dp.rate, dp.seed, drop_site(i, which), dp.path[i][path_ix], g.L)

    mk = branch_mask(g.depth - 1, SITE_DROP2, 1)
    dx = ops.layernorm_bwd(d, S["xf"], P["norm.weight"], S["meanf"], S["rstdf"], G["norm.weight"], G["norm.bias"], drop=mk)
    dx, dx_masked = dx if mk is not None else (dx, None)
    ready(["norm.weight", "norm.bias"])
    for i in range(g.depth - 1, -1, -1):
        b = f"blocks.{i}."
        s = S["blocks"][i]
        if "qkv" not in s:                       # checkpointed block: re-run its forward from the saved input
            _, s = _block_forward(g, P, Wc, i, s["x"])
        branch_drop = dp is not None and dp.branch_active(i)
        # gradient entering the MLP branch = the stream gradient through drop_path2 / drop2 (same mask, same scale)
        dbr = dx_masked if branch_drop else dx
        wgrad(dbr, s["h"], b + "mlp.fc2")
        drop1 = (dp.rate, dp.seed, drop_site(i, SITE_DROP1), None, 0) if (dp is not None and dp.rate > 0) else None
        fused = act == torch.bfloat16
        dpre = dgrad(dbr, Wc[b + "mlp.fc2.weight"], g.hidden, epi=EPI_DGELU, aux=s["pre"], drop=drop1 if fused else None)
        if drop1 is not None and not fused:
            ops.dropout(dpre, dp.rate, dp.seed, drop_site(i, SITE_DROP1), out=dpre)
        del dbr
        wgrad(dpre, s["y2"], b + "mlp.fc1")
        dy2 = dgrad(dpre, Wc[b + "mlp.fc1.weight"], D)
        del dpre
        mk = branch_mask(i, SITE_PROJ, 0)
        dxm = ops.layernorm_bwd(dy2, s["xm"], P[b + "norm2.weight"], s["mean2"], s["rstd2"], G[b + "norm2.weight"],
                                G[b + "norm2.bias"], dres=dx, drop=mk)
        dxm, dbr = dxm if mk is not None else (dxm, dxm)
        del dx_masked
        wgrad(dbr, s["ao"], b + "attn.proj")
        dao = dgrad(dbr, Wc[b + "attn.proj.weight"], D)
        del dbr
        adrop = (dp.rate, dp.seed, drop_site(i, SITE_ATTN)) if (dp is not None and dp.rate > 0) else None
        dqkv = ops.attn_bwd(s["qkv"], s["ao"], dao, s["lse"], g.B, g.L, g.heads, g.hd, adrop)
        wgrad(dqkv, s["y1"], b + "attn.qkv")
        dy1 = dgrad(dqkv, Wc[b + "attn.qkv.weight"], D)
        del dqkv
        mk = branch_mask(i - 1, SITE_DROP2, 1)
        if i == 0 and dp is not None and dp.rate > 0:       # the gradient leaving Block 0 goes through pos_drop: only the
            mk = (dp.rate, dp.seed, SITE_POS, None, 0)      # masked copy is needed below
        dx = ops.layernorm_bwd(dy1, s["x"], P[b + "norm1.weight"], s["mean1"], s["rstd1"], G[b + "norm1.weight"],
                               G[b + "norm1.bias"], dres=dxm, drop=mk)
        dx, dx_masked = dx if mk is not None else (dx, None)
        s.clear()
        ready([b + n for n in ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight",
                               "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias",
                               "mlp.fc2.weight", "mlp.fc2.bias")])
    # tok = pos_drop(proj(o) + bias + posres)
    if dp is not None and dp.rate > 0:
        dx = dx_masked if g.depth > 0 else ops.dropout(dx, dp.rate, dp.seed, SITE_POS, out=dx)
    dposres = torch.zeros(g.L * D, device=dev, dtype=torch.float32)
    ops.colsum(dx.view(g.B, g.L * D), dposres)
    wgrad(dx, S["o"], "var_agg.proj")
    ready(["var_agg.proj.weight", "var_agg.proj.bias"])
    do = dgrad(dx, Wc["var_agg.proj.weight"], D)
    dtab_s, dtab_v = ops.frontend_bwd(x, tab_s, tab_v, do, g.p, g.gh, g.gw, g.hd)
    return dtab_s, dtab_v, dposres.view(g.L, D)


def _wgrad_split(m: int, n: int, sms: int = 148, max_split: int = 8) -> int:
    """Smallest split-K factor whose work items (128 x 256 output tiles x splits) fill >= 85 % of the last wave."""
    tiles = ((m + 127) // 128) * ((n + 255) // 256)
    best, best_eff = 1, 0.0
    for sp in range(1, max_split + 1):
        items = tiles * sp
        eff = items / (((items + sms - 1) // sms) * sms)
        if eff >= 0.85:
            return sp
        if eff > best_eff:
            best, best_eff = sp, eff
    return best


class _BicubicResample(torch.autograd.Function):
    """pos_embed [ih*iw, D] -> [oh*ow, D] on the o2_bicubic kernels (channels-last: none of the reference's permutes)."""

    @staticmethod
    def forward(ctx, table, ih, iw, oh, ow):
        ctx.dims = (ih, iw, oh, ow)
        return ops.bicubic_fwd(table.detach(), ih, iw, oh, ow)

    @staticmethod
    def backward(ctx, d):
        return ops.bicubic_bwd(d.contiguous().float(), *ctx.dims), None, None, None, None


class ReslimFunction(torch.autograd.Function):
    """One autograd node for the whole kernel schedule: inputs (x, tab_s, tab_v, posres, *kernel params)."""

    @staticmethod
    def forward(ctx, g: Geometry, names: List[str], Wc_in: Optional[dict], x, tab_s, tab_v, posres, *params):
        lowp = g.act != torch.float32
        P = {}
        for n, p in zip(names, params):
            p = p.detach()
            is_gemm_w = n.endswith(".weight") and p.dim() == 2
            P[n] = p if (is_gemm_w and lowp) else _f32(p)        # big GEMM weights are never up-cast
        if Wc_in is not None:                  # operands maintained by a training engine (bf16 copy / sharded units)
            Wc = Wc_in
        elif not lowp:
            Wc = P
        else:
            Wc = {n: (P[n] if P[n].dtype == g.act else ops.cast_bf16(P[n].contiguous()))
                  for n in names if n.endswith(".weight") and P[n].dim() == 2}
        posres = posres.detach()
        posres_act = posres if not lowp else ops.cast_bf16(posres.contiguous())
        preds, S = reslim_forward(g, P, Wc, x, tab_s.detach(), tab_v.detach(), posres_act)
        ctx.g, ctx.names, ctx.P, ctx.Wc, ctx.S = g, names, P, Wc, S
        ctx.save_for_backward(x, tab_s, tab_v)
        ctx.param_dtypes = [p.dtype for p in params]
        return preds

    @staticmethod
    def backward(ctx, dpreds):
        g, names, P, Wc, S = ctx.g, ctx.names, ctx.P, ctx.Wc, ctx.S
        x, tab_s, tab_v = ctx.saved_tensors
        G = {}
        for n in names:
            G[n] = torch.zeros(P[n].shape, device=x.device, dtype=torch.float32)
        dpreds = dpreds.contiguous()
        if dpreds.dtype != g.act:
            dpreds = dpreds.to(g.act)
        dts, dtv, dpos = reslim_backward(g, P, Wc, x, tab_s.detach(), tab_v.detach(), S, dpreds, G)
        ctx.S = None
        grads = [G[n] if dt_ == torch.float32 else G[n].to(dt_) for n, dt_ in zip(names, ctx.param_dtypes)]
        return (None, None, None, None, dts, dtv, dpos, *grads)


# ------------------------------------------------------------------------------------------------
# the module
# ------------------------------------------------------------------------------------------------
class Res_Slim_ViT(nn.Module):
    """Same constructor as the reference (res_slimvit.py:22-43).  ``compute_dtype`` (extra, keyword-only) selects the
    activation / tensor-core operand type: torch.float32 (SIMT kernels, exact-parity arm) or torch.bfloat16 (tcgen05).
    If left None it follows the dtype of the parameters (``model.to(torch.bfloat16)`` == FSDP MixedPrecision bf16,
    intermediate_downscaling.py:601-607)."""

    def __init__(self, default_vars, img_size, in_channels, out_channels, history, superres_mag=4, cnn_ratio=4,
                 patch_size=16, drop_path=0.1, drop_rate=0.1, learn_pos_emb=False, embed_dim=1024, depth=24,
                 decoder_depth=8, num_heads=16, mlp_ratio=4.0, tensor_par_size=1, tensor_par_group=None,
                 FusedAttn_option=FusedAttn.CK, *, compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        if tensor_par_size != 1:
            raise NotImplementedError("orbit2_b200: tensor_par_size must be 1 (one 8-GPU box uses DP / FSDP sharding)")
        self.default_vars = list(default_vars)
        self.img_size = tuple(img_size)
        self.cnn_ratio = cnn_ratio
        self.superres_mag = superres_mag
        self.in_channels = in_channels * history
        self.out_channels = out_channels
        self.patch_size = patch_size
        self.history = history
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.depth = depth
        self.decoder_depth = decoder_depth
        self.spatial_resolution = 0
        self.tensor_par_size = tensor_par_size
        self.tensor_par_group = tensor_par_group
        self.drop_rate, self.drop_path = float(drop_rate), float(drop_path)
        self.compute_dtype = compute_dtype
        # per-Block activation recomputation (the reference applies checkpoint wrappers to every Block under FSDP,
        # intermediate_downscaling.py:583-590, 635-637): set True to keep only each block's input for backward
        self.activation_checkpointing = False
        assert embed_dim % num_heads == 0

        self.spatial_embed = nn.Linear(1, embed_dim)
        self.token_embeds = nn.ModuleList([_PatchEmbed(self.img_size, patch_size, embed_dim) for _ in self.default_vars])
        self.num_patches = self.token_embeds[0].num_patches
        self.var_embed = nn.Parameter(torch.zeros(1, len(self.default_vars), embed_dim), requires_grad=True)
        self.var_map = {v: i for i, v in enumerate(self.default_vars)}
        self.var_query = nn.Parameter(torch.zeros(1, 1, embed_dim), requires_grad=True)
        self.var_agg = _VarAgg(embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, embed_dim), requires_grad=learn_pos_emb)
        hidden = int(embed_dim * mlp_ratio)
        self.blocks = nn.ModuleList([Block(embed_dim, hidden) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim)
        self.path2 = nn.Sequential(
            nn.Conv2d(out_channels + 4, cnn_ratio * superres_mag * superres_mag, 3, 1, 1), nn.GELU(),
            nn.PixelShuffle(superres_mag), nn.Conv2d(cnn_ratio, out_channels, 3, 1, 1))
        head = []
        for _ in range(decoder_depth):
            head += [nn.Linear(embed_dim, embed_dim), nn.GELU()]
        head.append(nn.Linear(embed_dim, out_channels * (superres_mag * patch_size) ** 2))
        self.head = nn.Sequential(*head)
        self.conv_out = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        self.initialize_weights()
        self._names = kernel_param_names(depth, decoder_depth)
        self._wc_cache = None           # (versions, dict) bf16 copies of the GEMM weights
        self.external_wc = None         # set by TrainEngine.__init__: the operands it keeps fresh (flat bf16 buffer
                                        # refreshed by the fused AdamW, or the FULL_SHARD gather-on-lookup mapping)
        self.external_begin = None      # FULL_SHARD: called before every module forward (restarts the unit prefetch)
        self._warned_drop = False

    # res_slimvit.py:125-145
    def initialize_weights(self):
        pe = get_2d_sincos_pos_embed(self.embed_dim, self.img_size[0] // self.patch_size, self.img_size[1] // self.patch_size)
        self.pos_embed.data.copy_(torch.from_numpy(pe).float().unsqueeze(0))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    # res_slimvit.py:148-164
    def data_config(self, res, img_size, in_channels, out_channels):
        self.spatial_resolution = res
        self.img_size = tuple(img_size)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_patches = img_size[0] * img_size[1] // (self.patch_size ** 2)
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_rank() == 0:
            print("updated res is ", res, "img_size", img_size, "in_channels", in_channels, "out_channels", out_channels,
                  "num_patches", self.num_patches, flush=True)

    # res_slimvit.py:302-310
    def find_var_index(self, in_variables, out_variables):
        in_variables = list(in_variables)
        return [in_variables.index(v) for v in out_variables] + [in_variables.index(v) for v in STATIC_VARS]

    def _dev_const(self, key, values, dtype):
        """Small host constants of the parameter prep, uploaded once per (key, device): no host->device copy on the
        per-step path (also what lets a CUDA-graph capture of the step run through the prep)."""
        dev = self.var_embed.device
        cache = self.__dict__.setdefault("_dev_consts", {})
        t = cache.get((key, dev))
        if t is None:
            t = cache[(key, dev)] = torch.tensor(values, dtype=dtype, device=dev)
        return t

    def get_var_ids(self, variables):
        return [self.var_map[v] for v in variables]        # KeyError for unknown variables, like the reference

    # ------------------------------------------------------------------ host prep (parameters only)
    def frontend_tables(self, var_ids: Sequence[int]):
        """SURVEY.md appendix B.  tab_s [V, heads, PP+1] (scores, pre-scaled), tab_v [heads, V*(PP+1), hd]."""
        D, heads = self.embed_dim, self.num_heads
        hd = D // heads
        PP = self.patch_size ** 2
        V = len(var_ids)
        f = torch.float32
        Wt = torch.stack([self.token_embeds[i].proj.weight.to(f).reshape(D, PP) for i in var_ids])        # V,D,PP
        c = torch.stack([self.token_embeds[i].proj.bias.to(f) for i in var_ids]) + \
            self.var_embed.to(f)[0].index_select(0, self._dev_const(("ids", tuple(var_ids)), list(var_ids), torch.long))
        Wp = torch.cat([Wt, c.unsqueeze(-1)], dim=-1)                                                       # V,D,PP+1
        q = (self.var_agg.q.weight.to(f) @ self.var_query.to(f)[0, 0]).reshape(heads, hd) * hd ** -0.5
        Wk, Wv = self.var_agg.kv.weight.to(f)[:D].reshape(heads, hd, D), self.var_agg.kv.weight.to(f)[D:]
        qt = torch.einsum("he,hed->hd", q, Wk)                                                              # heads,D
        tab_s = torch.einsum("hd,vdk->vhk", qt, Wp).contiguous()
        M = torch.einsum("nd,vdk->vkn", Wv, Wp)                                                             # V,PP+1,D
        tab_v = M.reshape(V * (PP + 1), heads, hd).permute(1, 0, 2).contiguous()
        return tab_s, tab_v

    def pos_res_embed(self, gh: int, gw: int, act: torch.dtype):
        """pos_embed (+ on-the-fly bicubic resample, pos_embed.py:103-138) + spatial_embed(resolution) -> [L, D]."""
        pe = self.pos_embed.to(torch.float32)
        n = pe.shape[1]
        oh = int((n // 2) ** 0.5)
        if oh != gh:                                           # the reference assumes W/H == 2 for the stored grid
            if pe.shape[0] != 1 or n != 2 * oh * oh:
                raise RuntimeError(f"pos_embed with {n} rows is not an h x 2h grid (pos_embed.py:108-113 assumes W/H == 2)")
            pe = _BicubicResample.apply(pe[0].contiguous(), oh, 2 * oh, gh, gw).unsqueeze(0)
        res = self._dev_const(("res", float(self.spatial_resolution)), [float(self.spatial_resolution)], torch.float32)
        se = F.linear(res, self.spatial_embed.weight.to(torch.float32), self.spatial_embed.bias.to(torch.float32))
        return (pe[0] + se[None]).contiguous()           # fp32; the kernel schedule casts it to the activation dtype

    def _act_dtype(self):
        if self.compute_dtype is not None:
            return self.compute_dtype
        return self.blocks[0].attn.qkv.weight.dtype if self.depth else self.var_query.dtype

    def _param_dict(self):
        sd = dict(self.named_parameters())
        return sd

    def gemm_weights(self, act, params: Dict[str, torch.Tensor]):
        if self.external_wc is not None:
            if self.external_begin is not None:
                self.external_begin()
            return self.external_wc
        if act == torch.float32:
            return None
        names = [n for n in self._names if n.endswith(".weight") and params[n].dim() == 2]
        vers = tuple(params[n]._version for n in names) + tuple(params[n].data_ptr() for n in names)
        if self._wc_cache is not None and self._wc_cache[0] == vers:
            return self._wc_cache[1]
        wc = {}
        for n in names:
            w = params[n].detach()
            wc[n] = w.contiguous() if w.dtype == torch.bfloat16 else ops.cast_bf16(w.contiguous())
        self._wc_cache = (vers, wc)
        return wc

    def geometry(self, x, in_variables, out_variables, act) -> Geometry:
        g = Geometry()
        B, V, Hx, Wx = x.shape
        p = self.patch_size
        g.B, g.V, g.Hx, g.Wx, g.p = B, V, Hx, Wx, p
        H, W = self.img_size
        if (H, W) != (Hx, Wx):
            raise ValueError(f"input grid {Hx}x{Wx} differs from model.img_size {H}x{W}: call data_config() first")
        # the reference's unpatchify needs H*mag divisible by p (res_slimvit.py:174 raises on 181 rows)
        if (H * self.superres_mag) % p or (W * self.superres_mag) % p or H % p or W % p:
            raise RuntimeError(f"shape invalid: img_size {H}x{W} is not divisible by patch_size {p} (the reference's "
                               "unpatchify raises here too; crop 181-row ERA5 fields to 180 rows)")
        g.gh, g.gw = H // p, W // p
        g.L = g.gh * g.gw
        g.T = B * g.L
        g.D, g.heads = self.embed_dim, self.num_heads
        g.hd = g.D // g.heads
        g.depth, g.dec, g.mag = self.depth, self.decoder_depth, self.superres_mag
        g.C, g.cr = len(out_variables), self.cnn_ratio
        g.hidden = self.blocks[0].mlp.fc1.out_features if self.depth else 0
        g.idx7 = self.find_var_index(in_variables, out_variables)
        g.act = act
        g.ckpt = bool(getattr(self, "activation_checkpointing", False))
        g.drop = None
        static_plan = getattr(self, "static_drop_plan", None)     # CUDA-graph engines: one plan with fixed device buffers
        if self.training and static_plan is not None:
            if static_plan.B != B:
                raise RuntimeError("static_drop_plan was built for another batch size")
            g.drop = static_plan
        elif self.training and (self.drop_rate > 0 or self.drop_path > 0):
            # seeds come from torch's CPU generator: reproducible under torch.manual_seed, no device sync
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            dpr = [float(v) for v in torch.linspace(0, self.drop_path, self.depth)] if self.depth else []
            g.drop = DropPlan(self.drop_rate, dpr, B, seed, x.device)
        if g.C != self.conv_out.weight.shape[0]:
            raise ValueError(f"{g.C} output variables but the model was built for {self.conv_out.weight.shape[0]}")
        return g

    # res_slimvit.py:312-338
    def forward(self, x, in_variables, out_variables):
        if x.dim() == 5:
            x = x.flatten(1, 2)
        if not x.is_cuda:
            raise RuntimeError("orbit2_b200.Res_Slim_ViT runs on sm_100a CUDA kernels only (no CPU fallback)")
        act = self._act_dtype()
        x = x.contiguous().float()
        in_variables, out_variables = list(in_variables), list(out_variables)
        g = self.geometry(x, in_variables, out_variables, act)
        var_ids = self.get_var_ids(in_variables)
        params = self._param_dict()
        tab_s, tab_v = self.frontend_tables(var_ids)
        posres = self.pos_res_embed(g.gh, g.gw, act)
        wc = self.gemm_weights(act, params)
        return ReslimFunction.apply(g, self._names, wc, x, tab_s, tab_v, posres, *[params[n] for n in self._names])
