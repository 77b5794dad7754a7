"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference (ORBIT-2) on CPU.

Only usable where ``/root/reference`` (or ``baseline/_ref``) exists, i.e. in the build
container.  It is used by ``oracle/make_golden.py`` (to produce ``tests/golden/*.npz``) and by
the ``-m "not gpu"`` tests that pin ``oracle/reslim_oracle.py`` against the live reference.
Nothing in the product package imports this file.

The reference cannot be imported as a package in this image: ``climate_learn/__init__`` pulls
mpi4py / matplotlib / lpips and the model imports timm + xformers, none of which are
installed.  We pre-seed ``sys.modules`` with minimal stand-ins for the *third-party* symbols
 and import the reference's own files unchanged.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

_REF_CANDIDATES = ["/root/reference", os.path.join(os.path.dirname(__file__), "..", "baseline", "_ref")]


def reference_root() -> str | None:
    for c in _REF_CANDIDATES:
        if os.path.isdir(os.path.join(c, "src", "climate_learn")):
            return os.path.abspath(c)
    return None


class _DropPath(nn.Module):
    """timm.layers.DropPath (stochastic depth, scale_by_keep=True) -- third-party stand-in."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def _mod(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


_loaded = {}


def load_reference():
    """Returns a namespace with Res_Slim_ViT, FusedAttn, functional (losses), metrics."""
    if _loaded:
        return _loaded["ns"]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not available (expected /root/reference)")
    for n in ["timm", "timm.models", "timm.models.vision_transformer", "timm.layers", "timm.layers.helpers",
              "timm.layers.trace_utils", "timm.layers.grn", "xformers", "xformers.components",
              "xformers.components.attention", "xformers.components.attention.core"]:
        if n not in sys.modules:
            _mod(n)
    sys.modules["timm.models.vision_transformer"].trunc_normal_ = nn.init.trunc_normal_
    sys.modules["timm.layers"].DropPath = _DropPath
    sys.modules["timm.layers.helpers"].to_2tuple = lambda x: x if isinstance(x, (tuple, list)) else (x, x)
    sys.modules["timm.layers.trace_utils"]._assert = lambda c, m: None
    sys.modules["timm.layers.grn"].GlobalResponseNorm = nn.Identity
    sys.modules["xformers.components.attention.core"].scaled_dot_product_attention = None
    # losses: lpips / torchmetrics are import-only for the functions we use
    if "lpips" not in sys.modules:
        lp = _mod("lpips")
        lp.LPIPS = type("LPIPS", (nn.Module,), {})
        lp.NetLinLayer = type("NetLinLayer", (nn.Module,), {})
    if "torchmetrics" not in sys.modules:
        _mod("torchmetrics")
        _mod("torchmetrics.functional")
        tmi = _mod("torchmetrics.functional.image")
        tmi.image_gradients = None
    R = os.path.join(root, "src", "climate_learn")
    for n, p in [("climate_learn", R), ("climate_learn.utils", R + "/utils"), ("climate_learn.models", R + "/models"),
                 ("climate_learn.models.hub", R + "/models/hub"),
                 ("climate_learn.models.hub.components", R + "/models/hub/components"),
                 ("climate_learn.metrics", R + "/metrics")]:
        if n not in sys.modules:
            _mod(n).__path__ = [p]
    rs = importlib.import_module("climate_learn.models.hub.res_slimvit")
    fa = importlib.import_module("climate_learn.utils.fused_attn")
    fn = importlib.import_module("climate_learn.metrics.functional")
    ns = types.SimpleNamespace(Res_Slim_ViT=rs.Res_Slim_ViT, FusedAttn=fa.FusedAttn, functional=fn,
                               module=rs, root=root)
    try:
        ns.metrics = importlib.import_module("climate_learn.metrics.metrics")
    except Exception as e:  # torchvision vgg etc. -- not needed for the functional oracle
        ns.metrics = None
        ns.metrics_import_error = repr(e)
    _loaded["ns"] = ns
    return ns


# examples/intermediate_downscaling.py:267-278 cannot be imported (pulls the whole package), and
# it is 10 lines; era5_constants.py:83 gives CONSTANTS.  The oracle restates it.


def load_reference_loaders():
    """Imports the reference's ``climate_learn.utils.loaders`` UNMODIFIED (``load_architecture`` :259-378, ``load_loss``
    :436-450 -- the construction seam of the drop-in) on top of ``load_reference()``.  Its package-level
    imports pull the whole library, so the sub-packages are assembled by hand: the reference's own files are imported from
    where they lie, only the data package (mpi4py / xarray readers) is a stand-in exposing the one name ``loaders`` needs."""
    if "loaders" in _loaded:
        return _loaded["loaders"]
    ns = load_reference()
    R = os.path.join(ns.root, "src", "climate_learn")
    # third-party names the other hub models import (timm ViT blocks); never instantiated here
    vt = sys.modules["timm.models.vision_transformer"]
    for n in ("Block", "PatchEmbed"):
        if not hasattr(vt, n):
            setattr(vt, n, type(n, (nn.Module,), {}))
    data = sys.modules.get("climate_learn.data") or _mod("climate_learn.data")
    data.__path__ = [R + "/data"]
    if not hasattr(data, "IterDataModule"):
        data.IterDataModule = type("IterDataModule", (), {})
    if "climate_learn.data.processing" not in sys.modules:
        _mod("climate_learn.data.processing").__path__ = [R + "/data/processing"]
    hub = sys.modules["climate_learn.models.hub"]
    for sub, names in (("utils", ["MODEL_REGISTRY"]), ("climatology", ["Climatology"]), ("interpolation", ["Interpolation"]),
                       ("linear_regression", ["LinearRegression"]), ("persistence", ["Persistence"]), ("resnet", ["ResNet"]),
                       ("unet", ["Unet"]), ("vit", ["VisionTransformer"]), ("res_slimvit", ["Res_Slim_ViT"])):
        m = importlib.import_module("climate_learn.models.hub." + sub)
        for n in names:
            setattr(hub, n, getattr(m, n))
    models = sys.modules["climate_learn.models"]
    models.MODEL_REGISTRY = hub.MODEL_REGISTRY
    models.hub = hub
    importlib.import_module("climate_learn.models.lr_scheduler")
    if "climate_learn.transforms" not in sys.modules:
        _mod("climate_learn.transforms").__path__ = [R + "/transforms"]
    tr = sys.modules["climate_learn.transforms"]
    tr.TRANSFORMS_REGISTRY = importlib.import_module("climate_learn.transforms.registry").TRANSFORMS_REGISTRY
    for sub in ("denormalize", "mask"):
        importlib.import_module("climate_learn.transforms." + sub)
    met = sys.modules["climate_learn.metrics"]
    mu = importlib.import_module("climate_learn.metrics.utils")
    met.MetricsMetaInfo, met.METRICS_REGISTRY = mu.MetricsMetaInfo, mu.METRICS_REGISTRY
    if ns.metrics is None:
        raise RuntimeError("reference metrics module not importable: " + getattr(ns, "metrics_import_error", "?"))
    loaders = importlib.import_module("climate_learn.utils.loaders")
    _loaded["loaders"] = loaders
    return loaders
