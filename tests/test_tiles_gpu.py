"""GPU: tiled inference through the drop-in module (tile grid != init grid, so the on-the-fly bicubic resample of
pos_embed is exercised, pos_embed.py:103-138) against the oracle run tile by tile and stitched the reference's way."""
import pytest
import torch

from tests.util import build_model, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_tiled_inference_vs_oracle(dtype, tol):
    from oracle import cases, reslim_oracle as O
    from orbit2_b200 import tiles
    cfg = cases.get_case("8m")                       # 32x64 field, pos_embed stored for 16x32 patches
    div, overlap = 2, 4                              # tiles 20 x 40 (10 x 20 patches)
    th, tw = tiles.check_tiling(32, 64, div, div, overlap, cfg["patch_size"])
    assert (th, tw) == (20, 40)
    sd = O.init_state_dict(cfg, seed=2)
    x, _ = O.synthetic_batch(cfg, 2, cfg["in_vars"], cfg["out_vars"], seed=2)
    m = build_model(cfg, sd, "cuda", dtype).eval()
    m.img_size = (th, tw)                            # what data_config() sets for the tiled dataset
    with torch.no_grad():
        ours = tiles.tiled_forward(m, x.cuda(), cfg["in_vars"], cfg["out_vars"], div, overlap)
    # oracle: every tile on its own, inner region copied (utils/visualize.py:125-311)
    tcfg = dict(cfg, img_size=(th, tw))
    sd64 = {k: v.double() for k, v in sd.items()}
    top, bottom, left, right = tiles.overlap_margins(overlap)
    ref = torch.empty(2, 3, 128, 256, dtype=torch.float64)
    for v in range(div):
        yi1, yi2, yt1, yt2 = tiles.axis_bounds(32, div, v, top, bottom)
        for h in range(div):
            xi1, xi2, xt1, xt2 = tiles.axis_bounds(64, div, h, left, right)
            p = O.forward(sd64, tcfg, x[:, :, yi1:yi2, xi1:xi2].double(), cfg["in_vars"], cfg["out_vars"])
            ref[:, :, 64 * v:64 * (v + 1), 128 * h:128 * (h + 1)] = p[:, :, 4 * yt1:4 * yt2, 4 * xt1:4 * xt2]
    assert ours.shape == ref.shape
    assert rel(ours, ref) < tol
