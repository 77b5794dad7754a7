"""TEST INFRASTRUCTURE ONLY -- restates the keep-mask function of orbit2_b200/csrc/dropout.cu in torch integer
arithmetic so that parity tests can hand the SAME masks to the float64 oracle (oracle/reslim_oracle.py ``masks``).
Semantics of the masks themselves (x * keep / (1 - p)) are the reference's nn.Dropout / timm DropPath
(res_slimvit.py:284, attention.py:75,81, mlp.py:65,68, vit_blocks.py:78-79)."""
from __future__ import annotations

import math

import torch

M32 = 0xFFFFFFFF


def lowbias32(x: torch.Tensor) -> torch.Tensor:
    x = x & M32
    x = x ^ (x >> 16)
    x = (x * 0x21F0AAAD) & M32
    x = x ^ (x >> 15)
    x = (x * 0x735A2D97) & M32
    x = x ^ (x >> 15)
    return x


def _lb(x: int) -> int:
    return int(lowbias32(torch.tensor([x], dtype=torch.int64))[0])


def site_key(seed: int, site: int) -> int:
    return _lb((seed & M32) ^ _lb((site & M32) ^ ((seed >> 32) & M32)))


def keep_mask(seed: int, site: int, n: int, p: float) -> torch.Tensor:
    """bool [n]: keep decision of elements 0..n-1 of a dropout site (n < 2^33 in tests: hi32 of the pair index is 0)."""
    e = torch.arange(n, dtype=torch.int64)
    pair = e >> 1
    h = lowbias32((pair & M32) ^ site_key(seed, site) ^ (((pair >> 32) * 0x9E3779B1) & M32))
    half = torch.where((e & 1) == 1, h >> 16, h & 0xFFFF)
    return half >= math.floor(p * 65536.0)


def scaled_mask(seed: int, site: int, shape, p: float, dtype=torch.float64) -> torch.Tensor:
    n = 1
    for d in shape:
        n *= d
    return (keep_mask(seed, site, n, p).to(dtype) / (1.0 - p)).reshape(shape)


def attn_scaled_mask(seed: int, site: int, B: int, heads: int, N: int, p: float, dtype=torch.float64) -> torch.Tensor:
    """[B, heads, N, N] pre-scaled keep mask of the attention-probability dropout (orbit2_b200/csrc/common.cuh): one byte
    of lowbias32(((q >> 1) * ceil(N / 2) + (k >> 1)) ^ key_bh) per element, byte index (q & 1) * 2 + (k & 1), keep iff
    byte >= floor(p * 256); kept values scaled by 256 / (256 - floor(p * 256))."""
    thr8 = math.floor(p * 256.0)
    n2 = (N + 1) >> 1
    q = torch.arange(N, dtype=torch.int64).view(N, 1)
    k = torch.arange(N, dtype=torch.int64).view(1, N)
    blk = ((q >> 1) * n2 + (k >> 1)) & M32
    sh = ((q & 1) * 2 + (k & 1)) * 8
    sk = site_key(seed, site)
    out = torch.empty(B, heads, N, N, dtype=dtype)
    for bh in range(B * heads):
        key_bh = _lb(sk ^ ((bh * 0x9E3779B1) & M32))
        h = lowbias32(blk ^ key_bh)
        keep = ((h >> sh) & 0xFF) >= thr8
        out[bh // heads, bh % heads] = keep.to(dtype) * (256.0 / (256.0 - thr8))
    return out
