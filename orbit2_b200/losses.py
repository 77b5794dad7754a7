"""Loss classes behind the reference's ``METRICS_REGISTRY`` protocol, running on ``o2_loss_fwd_bwd``.

Mirrors src/climate_learn/metrics/utils.py:22-31 (``register`` / ``METRICS_REGISTRY`` / ``MetricsMetaInfo``) and
metrics.py:23-40 (``Metric`` ctor ``(aggregate_only=False, metainfo=None)``), :204-231 ``Bayesian_TV``, :236-263 ``MSE``,
:271-289 ``MAE``, :295-316 ``LatWeightedMSE``.  Return convention: 0-dim tensor when ``aggregate_only`` else ``[C+1]``
(per-channel means, then the aggregate) -- examples/intermediate_downscaling.py:301-304 handles both.

Reference defects that are *not* reproduced (SURVEY.md headline 4): ``LatWeightedMSE.__call__`` raises TypeError in the
reference (positional-argument mix-up); here it computes what its docstring says (functional ``mse(...,
lat_weights=w)``).  ``MAE.__call__`` of the reference rejects the ``var_names/var_weights`` keywords the driver passes; here
they are accepted and ignored (``mae`` has no variable weights, functional.py:218-232).

Extension used by the fused training step: every loss accepts ``clip_out_variables=[...]``; when given, the
driver's ``clip_replace_constant`` (intermediate_downscaling.py:267-278) is applied inside the same kernel (clamp
precipitation at 0, constant fields copied from the target) and ``pred`` is the *raw* model output.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from ._lib import LOSS_BAYESIAN_TV, LOSS_MAE, LOSS_MSE

# src/climate_learn/data/processing/era5_constants.py:83
CONSTANTS = ["orography", "land_sea_mask", "slt", "lattitude", "longitude"]
PRECIP = "total_precipitation_24hr"

METRICS_REGISTRY: Dict[str, type] = {}


def register(name):
    def decorator(metric_class):
        METRICS_REGISTRY[name] = metric_class
        metric_class.name = name
        return metric_class
    return decorator


@dataclass
class MetricsMetaInfo:
    in_vars: List[str]
    out_vars: List[str]
    lat: object
    lon: object
    climatology: object = None


def clip_spec(out_variables: Sequence[str]):
    """(clamp channel, constant-channel bit mask) of clip_replace_constant; raises ValueError like the reference
    (``list.index``) when precipitation is not an output variable."""
    out_variables = list(out_variables)
    clamp = out_variables.index(PRECIP)
    mask = 0
    for i, v in enumerate(out_variables):
        if v in CONSTANTS:
            mask |= 1 << i
    return clamp, mask


def clip_replace_constant(y, yhat, out_variables):
    """Drop-in for intermediate_downscaling.py:267-278 / utils/visualize.py:23-34 (in place on ``yhat``, no autograd:
    evaluation path).  Training uses the fused form (``clip_out_variables=`` on the loss)."""
    clamp, mask = clip_spec(out_variables)
    with torch.no_grad():
        ops.clip_replace_(yhat, y.float().contiguous(), clamp, mask)
    return yhat


class _LossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, kind, lat_w, ch_w, clamp_ch, const_mask):
        want = pred.requires_grad
        vec, dpred = ops.loss_fwd_bwd(pred, target, kind, lat_w=lat_w, ch_w=ch_w, clamp_ch=clamp_ch,
                                      const_mask=const_mask, want_grad=want)
        ctx.dpred = dpred
        return vec

    @staticmethod
    def backward(ctx, gvec):
        dpred = ctx.dpred
        ctx.dpred = None
        C = gvec.numel() - 1
        # aggregate = mean_c(per_channel[c])  =>  d/dpred = (g_agg + C * g_c) * d(aggregate)/dpred on channel c
        scale = (gvec[-1] + C * gvec[:-1]).float().contiguous()
        ops.scale_channels_(dpred, scale)
        return dpred, None, None, None, None, None, None


class Metric:
    kind = LOSS_MSE
    uses_var_weights = True

    def __init__(self, aggregate_only: bool = False, metainfo: Optional[MetricsMetaInfo] = None):
        self.aggregate_only = aggregate_only
        self.metainfo = metainfo
        self._lat_dev = None
        self._chw_cache = {}

    def _lat(self, pred):
        return None

    def _ch_w(self, pred, var_names, var_weights):
        if not self.uses_var_weights or var_names is None:
            return None
        if len(var_names) != pred.shape[1]:
            raise ValueError(f"{len(var_names)} variable names for {pred.shape[1]} channels")
        var_weights = var_weights or {}
        key = (tuple(var_names), tuple(sorted(var_weights.items())), pred.device)
        w = self._chw_cache.get(key)
        if w is None:
            w = torch.tensor([float(var_weights.get(v, 1.0)) for v in var_names], dtype=torch.float32, device=pred.device)
            self._chw_cache[key] = w
        return w

    def vector(self, pred, target, var_names=None, var_weights=None, clip_out_variables=None):
        """[C+1] loss vector (autograd-connected to ``pred``)."""
        if not pred.is_cuda:
            raise RuntimeError("orbit2_b200 losses run on sm_100a CUDA kernels only (no CPU fallback)")
        clamp, mask = (-1, 0) if clip_out_variables is None else clip_spec(clip_out_variables)
        if pred.dtype not in (torch.float32, torch.bfloat16):
            pred = pred.float()
        target = target.float()
        if target.shape[2] < pred.shape[2] or target.shape[3] < pred.shape[3]:
            raise ValueError(f"target {tuple(target.shape)} smaller than prediction {tuple(pred.shape)}")
        return _LossFunction.apply(pred.contiguous(), target.contiguous(), self.kind, self._lat(pred),
                                   self._ch_w(pred, var_names, var_weights), clamp, mask)

    def __call__(self, pred, target, var_names=None, var_weights=None, clip_out_variables=None):
        v = self.vector(pred, target, var_names, var_weights, clip_out_variables)
        return v[-1] if self.aggregate_only else v


class LatitudeWeightedMetric(Metric):
    """metrics.py:55-75: w = cos(lat) / mean(cos(lat)), shape [1,1,H,1] float64 (kept as ``lat_weights``)."""

    def __init__(self, aggregate_only: bool = False, metainfo: Optional[MetricsMetaInfo] = None):
        super().__init__(aggregate_only, metainfo)
        w = np.cos(np.deg2rad(np.asarray(self.metainfo.lat, dtype=np.float64)))
        w = w / w.mean()
        self.lat_weights = torch.from_numpy(w).view(1, 1, -1, 1)

    def cast_to_device(self, pred):
        self.lat_weights = self.lat_weights.to(device=pred.device)

    def _lat(self, pred):
        if self._lat_dev is None or self._lat_dev.device != pred.device:
            self._lat_dev = self.lat_weights.reshape(-1).to(device=pred.device, dtype=torch.float32).contiguous()
        if self._lat_dev.numel() != pred.shape[2]:
            raise ValueError(f"{self._lat_dev.numel()} latitude weights for a prediction with {pred.shape[2]} rows")
        return self._lat_dev


@register("mse")
class MSE(Metric):
    kind = LOSS_MSE


@register("bayesian_tv")
class Bayesian_TV(Metric):
    kind = LOSS_BAYESIAN_TV


@register("mae")
class MAE(Metric):
    kind = LOSS_MAE
    uses_var_weights = False


@register("lat_mse")
class LatWeightedMSE(LatitudeWeightedMetric):
    kind = LOSS_MSE
    uses_var_weights = False


@register("lat_bayesian_tv")
class LatWeightedBayesianTV(LatitudeWeightedMetric):
    """Not in the reference registry: functional bayesian_tv(..., lat_weights=w) (functional.py:117-167) as a class."""
    kind = LOSS_BAYESIAN_TV


@register("lat_mae")
class LatWeightedMAE(LatitudeWeightedMetric):
    """Not in the reference registry: functional mae(..., lat_weights=w) (functional.py:218-232) as a class."""
    kind = LOSS_MAE
    uses_var_weights = False


# ------------------------------------------------------------------------------------------------ evaluation metrics
class _EvalMetric:
    """Validation / test metrics of the downscaling task (utils/loaders.py:251-252: rmse, pearson, mean_bias) on ONE pass
    over prediction and target (``o2_eval_stats``).  ``denorm=(scale[C], shift[C])`` folds the reference's
    ``TransformedMetric(denormalise, metric)`` (metrics/metrics.py:100-115) into the same pass.  Evaluation only: the
    result carries no autograd history."""

    def __init__(self, aggregate_only: bool = False, metainfo: Optional[MetricsMetaInfo] = None, denorm=None):
        self.aggregate_only = aggregate_only
        self.metainfo = metainfo
        self.denorm = denorm
        self._dev = {}

    def _lat(self, pred):
        return None

    def _stats(self, pred, target):
        if not pred.is_cuda:
            raise RuntimeError("orbit2_b200 metrics run on sm_100a CUDA kernels only (no CPU fallback)")
        if pred.dtype not in (torch.float32, torch.bfloat16):
            pred = pred.float()
        sc = sh = None
        if self.denorm is not None:
            key = pred.device
            if key not in self._dev:
                self._dev[key] = tuple(torch.as_tensor(np.asarray(v, dtype=np.float32)).to(pred.device) for v in self.denorm)
            sc, sh = self._dev[key]
        with torch.no_grad():
            return ops.eval_stats(pred.detach().contiguous(), target.float().contiguous(), lat_w=self._lat(pred),
                                  scale=sc, shift=sh), pred.shape[2] * pred.shape[3]

    def _finish(self, per_channel):
        agg = per_channel.mean()
        out = agg if self.aggregate_only else torch.cat((per_channel, agg.unsqueeze(0)))
        return out.float()

    def _compose(self, s, n):
        raise NotImplementedError

    def __call__(self, pred, target, **_ignored):
        s, n = self._stats(pred, target)
        return self._finish(self._compose(s, n))


@register("rmse")
class RMSE(_EvalMetric):
    """functional.py:236-257: sqrt of the spatial mean per (sample, channel), then the batch mean."""

    def _compose(self, s, n):
        return (s[:, :, 0] / n).sqrt().mean(0)


@register("lat_rmse")
class LatWeightedRMSE(RMSE):
    def __init__(self, aggregate_only: bool = False, metainfo: Optional[MetricsMetaInfo] = None, denorm=None):
        super().__init__(aggregate_only, metainfo, denorm)
        w = np.cos(np.deg2rad(np.asarray(self.metainfo.lat, dtype=np.float64)))
        self.lat_weights = torch.from_numpy(w / w.mean()).view(1, 1, -1, 1)
        self._lat_dev = None

    def _lat(self, pred):
        if self._lat_dev is None or self._lat_dev.device != pred.device:
            self._lat_dev = self.lat_weights.reshape(-1).to(device=pred.device, dtype=torch.float32).contiguous()
        return self._lat_dev


@register("pearson")
class Pearson(_EvalMetric):
    """functional.py:294-309: cosine similarity of the mean-removed channel over all of B x H x W."""

    def _compose(self, s, n):
        t = s.sum(0)                                       # [C, 6]
        N = n * s.shape[0]
        sp, st_, spp, stt, spt = t[:, 1], t[:, 2], t[:, 3], t[:, 4], t[:, 5]
        cov = spt - sp * st_ / N
        vp = (spp - sp * sp / N).clamp_min(0).sqrt().clamp_min(1e-8)     # F.cosine_similarity clamps each norm at eps = 1e-8
        vt = (stt - st_ * st_ / N).clamp_min(0).sqrt().clamp_min(1e-8)
        return cov / (vp * vt)


@register("mean_bias")
class MeanBias(_EvalMetric):
    """functional.py:312-324: mean(target) - mean(pred) per channel."""

    def _compose(self, s, n):
        t = s.sum(0)
        return (t[:, 2] - t[:, 1]) / (n * s.shape[0])
