"""CPU: TILES geometry against a literal restatement of the reference's index arithmetic, and the halo exchange /
stitching on a world-size-4 gloo group (2 x 2 tiles)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orbit2_b200 import tiles


def reference_axis(n_in, n_out, div, idx, lo, hi):
    """src/climate_learn/data/iterdataset.py:123-170 + utils/visualize.py:125-233, one axis, literally
    (with the `**` typo of visualize.py:211 read as `*`)."""
    mul = n_out // n_in
    if div == 1:
        return 0, n_in, 0, n_out, 0, n_in, 0, n_out
    i1, i2 = n_in // div * idx, n_in // div * (idx + 1)
    o1, o2 = n_out // div * idx, n_out // div * (idx + 1)
    if idx == 0:
        i2 += lo; o2 += lo * mul
    else:
        i1 -= lo; o1 -= lo * mul
    if idx == div - 1:
        i1 -= hi; o1 -= hi * mul
    else:
        i2 += hi; o2 += hi * mul
    if idx == 0:
        i1t = 0; o1t = 0
    elif idx == div - 1:
        i1t = lo + hi; o1t = (lo + hi) * mul
    else:
        i1t = lo; o1t = lo * mul
    return i1, i2, o1, o2, i1t, i1t + n_in // div, o1t, o1t + n_out // div


@pytest.mark.parametrize("overlap", [0, 1, 2, 3, 4, 7])
@pytest.mark.parametrize("div", [1, 2, 3, 4])
def test_axis_bounds_match_reference(div, overlap):
    top, bottom, left, right = tiles.overlap_margins(overlap)
    if overlap % 2 == 0:
        assert (top, bottom, left, right) == (overlap // 2, overlap // 2, overlap // 2 * 2, overlap // 2 * 2)
    else:
        assert (top, bottom, left, right) == (overlap // 2, overlap // 2 + 1, overlap // 2 * 2, (overlap // 2 + 1) * 2)
    for n, lo, hi in ((48, top, bottom), (96, left, right)):
        sizes = set()
        for idx in range(div):
            i1, i2, t1, t2 = tiles.axis_bounds(n, div, idx, lo, hi)
            r = reference_axis(n, 4 * n, div, idx, lo, hi)
            assert (i1, i2, t1, t2) == (r[0], r[1], r[4], r[5])
            assert (4 * i1, 4 * i2, 4 * t1, 4 * t2) == (r[2], r[3], r[6], r[7])
            assert i1 + t1 == n // div * idx                       # the inner region is the tile's own block
            sizes.add(i2 - i1)
        assert len(sizes) == 1                                      # identical tile size (itermodule.py:170-175)
        assert sizes.pop() == (n if div == 1 else n // div + lo + hi)


def test_check_tiling():
    assert tiles.check_tiling(180, 360, 2, 2, 2, 2) == (92, 184)
    with pytest.raises(ValueError):
        tiles.check_tiling(180, 360, 2, 2, 1, 2)                    # 91 x 182 tile is not a whole number of 2x2 patches
    with pytest.raises(ValueError):
        tiles.check_tiling(181, 360, 2, 2, 2, 2)


class _Upsample(torch.nn.Module):
    """Stand-in 'network': 4x nearest upsampling of the first C channels plus a position-dependent term, so that both
    the tile content and the stitched position are checked."""
    superres_mag = 4

    def forward(self, x, in_vars, out_vars):
        y = torch.nn.functional.interpolate(x[:, :len(out_vars)], scale_factor=4, mode="nearest")
        return y + x[:, -1:].mean() * 0        # touches every input pixel, contributes nothing


def test_tiled_forward_single_device():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 24, 48, generator=g)
    m = _Upsample()
    full = m(x, None, ["a", "b", "c"])
    for div, ov in ((1, 0), (2, 2), (3, 3), (4, 1)):
        out = tiles.tiled_forward(m, x, None, ["a", "b", "c"], div, ov)
        assert torch.equal(out, full), (div, ov)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for (dv, dh, ov) in ((2, 2, 2), (2, 2, 3), (1, 4, 4), (4, 1, 1)):
        H, W = 24, 48
        g = torch.Generator().manual_seed(1)
        x = torch.randn(2, 5, H, W, generator=g)                        # every rank builds the same global field
        geo = tiles.ShardedField(H, W, dv, dh, ov)
        oy1, oy2, ox1, ox2 = geo.own(rank)
        tile = tiles.exchange_halos(x[:, :, oy1:oy2, ox1:ox2].contiguous(), geo, rank)
        ny1, ny2, nx1, nx2 = geo.need(rank)
        ok &= torch.equal(tile, x[:, :, ny1:ny2, nx1:nx2])              # == the reference's tile slice
        m = _Upsample()
        blk = tiles.sharded_tiled_forward(m, x[:, :, oy1:oy2, ox1:ox2].contiguous(), None, ["a", "b"], geo, rank)
        full = tiles.gather_output(blk, geo)
        ok &= torch.equal(full, m(x, None, ["a", "b"]))
        ok &= geo.halo_bytes(rank, 5, 2) == (tile.numel() - (oy2 - oy1) * (ox2 - ox1) * 10) * 4
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_halo_exchange_world4():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(4, _free_port(), ret), nprocs=4, join=True)
    assert dict(ret) == {0: True, 1: True, 2: True, 3: True}
