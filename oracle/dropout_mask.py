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


def step_word_mix(word) -> int:
    """Contribution of the device-resident step word (o2_dropout_seed_source, CUDA-graph replay) to every key:
    common.cuh step_word_mix; 0 when no word is installed.  ``word`` is the 64-bit pattern (negative int64 values wrap)."""
    if word is None:
        return 0
    w = int(word) & 0xFFFFFFFFFFFFFFFF
    return _lb((w & M32) ^ _lb(((w >> 32) & M32) ^ 0x5BD1E995))


def site_key(seed: int, site: int, step_word=None) -> int:
    return _lb((seed & M32) ^ _lb((site & M32) ^ ((seed >> 32) & M32))) ^ step_word_mix(step_word)


def keep_mask(seed: int, site: int, n: int, p: float, step_word=None) -> torch.Tensor:
    """bool [n]: keep decision of elements 0..n-1 of a dropout site (n < 2^33 in tests: hi32 of the pair index is 0)."""
    e = torch.arange(n, dtype=torch.int64)
    pair = e >> 1
    h = lowbias32((pair & M32) ^ site_key(seed, site, step_word) ^ (((pair >> 32) * 0x9E3779B1) & M32))
    half = torch.where((e & 1) == 1, h >> 16, h & 0xFFFF)
    return half >= math.floor(p * 65536.0)


def scaled_mask(seed: int, site: int, shape, p: float, dtype=torch.float64) -> torch.Tensor:
    n = 1
    for d in shape:
        n *= d
    return (keep_mask(seed, site, n, p).to(dtype) / (1.0 - p)).reshape(shape)


_KEEP_MUL = (0x9E3779B1, 0x85EBCA77, 0xC2B2AE3D, 0x27D4EB2F, 0x165667B1, 0xD3A2646D, 0xFD7046C5, 0xB55A4F09)


def attn_keep_bit(k: torch.Tensor) -> torch.Tensor:
    """bit of the 32-key keep word that belongs to key k (orbit2_b200/csrc/common.cuh attn_keep_bit)."""
    kk = k & 31
    return 7 - (kk >> 2) + 8 * (kk & 1) + 16 * ((kk >> 1) & 1)


def attn_scaled_mask(seed: int, site: int, B: int, heads: int, N: int, p: float, dtype=torch.float64,
                     step_word=None) -> torch.Tensor:
    """[B, heads, N, N] pre-scaled keep mask of the attention-probability dropout (orbit2_b200/csrc/common.cuh, bit-sliced):
    one 32-bit keep word per (query q, 32-key block kb): base = lowbias32((q * ceil(N / 32) + kb) ^ key_bh), eight planes
    w_i = lo32(base * K_i) ^ hi32(base * K_i) of a uniform byte U per key, keep iff U >= thr (LSB plane first:
    ge = w_i & ge where the threshold bit is set, w_i | ge where it is clear); key k reads bit attn_keep_bit(k).
    16-bit drop probability thr16 = floor(p * 65536) = 256 hi8 + frac8: the byte threshold of a word is hi8 + 1 when its
    dither byte lowbias32(base ^ 0x68E31DA4) >> 24 is below frac8, else hi8; kept values are scaled by
    65536 / (65536 - thr16)."""
    thr16 = math.floor(min(p, 0.99) * 65536.0) if p > 0 else 0
    hi8, frac8 = thr16 >> 8, thr16 & 0xFF
    nkb = (N + 31) >> 5
    q = torch.arange(N, dtype=torch.int64).view(N, 1)
    kb = torch.arange(nkb, dtype=torch.int64).view(1, nkb)
    k = torch.arange(N, dtype=torch.int64)
    bit = attn_keep_bit(k).view(1, N)
    sk = site_key(seed, site, step_word)
    out = torch.empty(B, heads, N, N, dtype=dtype)
    lo16 = 0xFFFF
    for bh in range(B * heads):
        key_bh = _lb(sk ^ ((bh * 0x9E3779B1) & M32))
        base = lowbias32(((q * nkb + kb) & M32) ^ key_bh)                      # [N, nkb], < 2^32
        ge = torch.full_like(base, M32)
        gh = torch.full_like(base, M32)
        for i, mul in enumerate(_KEEP_MUL):
            # 32 x 32 -> 64-bit product in int64 pieces (int64 would overflow on the full product)
            b_lo, b_hi = base & lo16, base >> 16
            m_lo, m_hi = mul & lo16, mul >> 16
            p0 = b_lo * m_lo
            p1 = b_lo * m_hi + b_hi * m_lo + (p0 >> 16)                         # < 2^33
            lo32 = ((p1 & lo16) << 16) | (p0 & lo16)
            hi32 = (b_hi * m_hi + (p1 >> 16)) & M32
            w = lo32 ^ hi32
            ge = (w & ge) if (hi8 >> i) & 1 else (w | ge)
            gh = (w & gh) if ((hi8 + 1) >> i) & 1 else (w | gh)
        word = torch.where((lowbias32(base ^ 0x68E31DA4) >> 24) < frac8, gh, ge)
        keep = (word[:, k >> 5] >> bit) & 1                                     # [N, N]
        out[bh // heads, bh % heads] = keep.to(dtype) * (65536.0 / (65536.0 - thr16))
    return out
