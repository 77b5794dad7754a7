"""Per-parameter relative gradient error of the bf16 (tcgen05) arm against the float64 fixtures / oracle, next to the error
of the REFERENCE SCHEDULE itself when it runs in bf16 (the oracle port under torch.autocast(bf16) on this GPU: fp32 master
weights, bf16 GEMM / attention operands, fp32 accumulation -- what intermediate_downscaling.py:601-607 does).  The second
column is the arithmetic's own floor for each parameter; tests/test_model_gpu.py bounds ours by it.

    python tools/bf16_grad_error.py > profiles/r02_bf16_grad_error.md"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import cases, reslim_oracle as O  # noqa: E402
from tests.test_model_gpu import run_ours  # noqa: E402
from tests.util import load_golden, rel  # noqa: E402


def oracle_autocast_grads(cfg, sd, x, y, loss_name, lat_w):
    sdc = {k: v.float().cuda().requires_grad_(True) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = O.training_step(sdc, cfg, x.float().cuda(), y.float().cuda(), cfg["in_vars"], cfg["out_vars"], loss_name,
                               cfg["var_weights"], lat_w.float().cuda() if lat_w is not None else None)
    loss.float().backward()
    return {k: v.grad.double().cpu() for k, v in sdc.items() if v.grad is not None}


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    print("# bf16 parameter-gradient error: this repo vs the reference schedule under bf16 autocast (both against float64)\n")
    for fixture in ("tiny_mse", "tiny_bayesian_tv_lat", "tiny_prism_mae_lat"):
        z, meta, sd, gref = load_golden(fixture)
        cfg = cases.get_case(meta[0])
        x, y = torch.from_numpy(z["x"]), torch.from_numpy(z["y"])
        use_lat = meta[4] == "1"
        _, _, grads = run_ours(cfg, sd, x, y, meta[2], use_lat, torch.bfloat16, z["lat"])
        lat_w = None
        if use_lat:
            lat_w = O.lat_weights(z["lat"])
        ref = oracle_autocast_grads(cfg, sd, x, y, meta[2], lat_w)
        print(f"## {fixture} ({meta[2]}{', latitude weighted' if use_lat else ''})\n")
        print("| parameter | ours bf16 | reference schedule bf16 autocast |")
        print("|---|---|---|")
        rows = []
        for k, g in gref.items():
            if g.abs().max() == 0 or k not in grads:
                continue
            rows.append((k, rel(grads[k], g), rel(ref[k], g) if k in ref else float("nan")))
        for k, a, b in sorted(rows, key=lambda r: -r[1])[:14]:
            print(f"| `{k}` | {a:.2e} | {b:.2e} |")
        print(f"| **max over {len(rows)} parameters** | **{max(r[1] for r in rows):.2e}** | **{max(r[2] for r in rows):.2e}** |")
        print(f"| median | {float(np.median([r[1] for r in rows])):.2e} | {float(np.median([r[2] for r in rows])):.2e} |\n")


if __name__ == "__main__":
    main()
