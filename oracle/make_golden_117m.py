"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/117m_fullgrid_f64_compact.npz: the float64 oracle (oracle/reslim_oracle.py,
itself pinned to the live reference by tests/test_oracle.py) run ONCE on BASELINE configs[1] at full size -- interm_117m,
180x360 -> 720x1440 (L = 16200), B = 1, seeded weights (init_state_dict(cfg, 0)) and batch (synthetic_batch(cfg, 1, seed 0)),
Bayesian-TV training loss -- which takes ~4.5 min and 33 GB on 8 cores, too slow for the GPU test run.  Stored compactly:
  loss                     float64 scalar
  pred_idx / pred_val      8192 sampled elements of the clipped prediction [1,3,720,1440] (+ pred_absmax)
  per parameter k:  g_idx/k, g_val/k = up to 2048 sampled gradient elements (float64), g_absmax/k = max |grad| over the
                    WHOLE gradient (the normaliser of the relative error), o32_err/k = the fp32 CPU oracle's own max
                    deviation from float64 relative to g_absmax (fp32 + SDPA, what tests can run live): a per-parameter
                    measure of how much fp32 arithmetic alone moves that gradient at this size.
  sd_checksum              sum of |w| over all weights (the seeded init must reproduce on the test box)
Run:  python -m oracle.make_golden_117m        (needs ~40 GB of host memory)"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import cases, reslim_oracle as O

N_GRAD, N_PRED = 2048, 8192


def sample_idx(n: int, k: int, seed: int) -> np.ndarray:
    g = torch.Generator().manual_seed(seed)
    return torch.randperm(n, generator=g)[:min(n, k)].sort().values.numpy()


def run(dtype):
    cfg = cases.get_case("117m")
    sd = O.init_state_dict(cfg, seed=0)
    x, y = O.synthetic_batch(cfg, 1, cfg["in_vars"], cfg["out_vars"], seed=0)
    sdr = {k: v.to(dtype).requires_grad_(True) for k, v in sd.items()}
    taps = {}
    loss = O.training_step(sdr, cfg, x.to(dtype), y.to(dtype), cfg["in_vars"], cfg["out_vars"], "bayesian_tv",
                           cfg["var_weights"], None, taps)
    loss.backward()
    return sd, loss.detach(), taps["preds"].detach(), {k: v.grad for k, v in sdr.items() if v.grad is not None}


def build():
    torch.set_num_threads(os.cpu_count() or 1)
    O.USE_SDPA = True
    sd, loss64, pred64, g64 = run(torch.float64)
    _, loss32, pred32, g32 = run(torch.float32)
    out = {"loss": np.float64(loss64.item()), "loss_f32_oracle": np.float64(loss32.item()),
           "sd_checksum": np.float64(sum(v.double().abs().sum().item() for v in sd.values()))}
    pi = sample_idx(pred64.numel(), N_PRED, 1)
    out["pred_idx"], out["pred_val"] = pi, pred64.reshape(-1).numpy()[pi]
    out["pred_absmax"] = np.float64(pred64.abs().max().item())
    out["pred_o32_err"] = np.float64((pred32.double() - pred64).abs().max().item() / pred64.abs().max().item())
    names = sorted(g64)
    for j, k in enumerate(names):
        g = g64[k].reshape(-1)
        amax = g.abs().max().item()
        if amax == 0:
            continue
        idx = sample_idx(g.numel(), N_GRAD, 100 + j)
        out["g_idx/" + k], out["g_val/" + k] = idx, g.numpy()[idx]
        out["g_absmax/" + k] = np.float64(amax)
        out["o32_err/" + k] = np.float64((g32[k].double().reshape(-1) - g).abs().max().item() / amax)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "117m_fullgrid_f64_compact.npz")
    res = build()
    np.savez_compressed(path, **res)
    worst = sorted(((float(v), k[8:]) for k, v in res.items() if k.startswith("o32_err/")), reverse=True)[:5]
    print("wrote", os.path.normpath(path), os.path.getsize(path) // 1024, "KiB; loss", float(res["loss"]),
          "fp32-oracle worst deviations", worst)
