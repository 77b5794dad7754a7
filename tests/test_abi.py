"""CPU: the C-ABI library loads without a GPU and exports every symbol include/o2b200.h declares; compute calls fail
loudly (no CPU fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "o2b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(o2_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from orbit2_b200 import _lib, build
    build.build_library()
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/o2b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert lib.o2_version() >= 100
    assert set(_lib.SIGNATURES) == set(syms)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from orbit2_b200 import _lib, losses
    from orbit2_b200.reslim import Res_Slim_ViT
    lib = _lib.load()
    assert lib.o2_device_ok() == 0
    with pytest.raises(_lib.O2Error):
        _lib.load(require_device=True)
    m = Res_Slim_ViT(["land_sea_mask", "orography", "lattitude", "landcover", "total_precipitation_24hr"], (4, 8), 5, 1, 1,
                     patch_size=2, embed_dim=64, depth=1, decoder_depth=1, num_heads=1)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 5, 4, 8), m.default_vars, ["total_precipitation_24hr"])
    with pytest.raises(RuntimeError):
        losses.MSE()(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))
