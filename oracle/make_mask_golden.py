"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/keep_masks.npz: known-answer vectors of the two dropout mask functions
(token-stream hash of csrc/dropout.cu, bit-sliced attention mask of csrc/common.cuh) as restated in oracle/dropout_mask.py,
packed to bits.  The GPU parity tests tie the kernels to the restatement; this fixture ties the restatement to a committed
answer, so neither side can drift silently.  Run:  python -m oracle.make_mask_golden"""
from __future__ import annotations

import os

import numpy as np

from . import dropout_mask as DM

SEED, SITE = 0x1234_5678_9ABC_DEF0, 17
CASES = {"attn_p10_n100": (0.1, 100), "attn_p35_n77": (0.35, 77)}


def build():
    out = {}
    for name, (p, n) in CASES.items():
        m = DM.attn_scaled_mask(SEED, SITE, 2, 3, n, p)
        out[name] = np.packbits((m > 0).numpy().reshape(-1))
    out["tokens_p10_n5000"] = np.packbits(DM.keep_mask(SEED, SITE, 5000, 0.1).numpy())
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "keep_masks.npz")
    np.savez_compressed(path, **build())
    print("wrote", os.path.normpath(path), {k: int(v.size) for k, v in build().items()})
