"""Data path: shard order, rank partition and TILES slices of orbit2_b200.data against the live
reference's NpyReader -> Downscale -> IndividualDataIter chain (CPU), the oracle's transform / metric restatements against
the live reference, and (GPU) the device-side normalisation + evaluation statistics against the oracle."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import reslim_oracle as O

IN_VARS = ["land_sea_mask", "orography", "lattitude", "landcover", "2m_temperature", "total_precipitation_24hr"]
OUT_VARS = ["total_precipitation_24hr", "2m_temperature"]


def make_shards(root, n_files=4, n_t=5, h=8, w=16, mag=4, seed=0):
    rng = np.random.default_rng(seed)
    inp, out = os.path.join(root, "lo"), os.path.join(root, "hi")
    for d, vars_, (hh, ww) in ((inp, IN_VARS, (h, w)), (out, OUT_VARS, (h * mag, w * mag))):
        for split in ("train", "val"):
            os.makedirs(os.path.join(d, split), exist_ok=True)
            for f in range(n_files if split == "train" else 1):
                arrs = {}
                for v in vars_:
                    a = rng.standard_normal((n_t, 1, hh, ww)).astype(np.float32) * 3 + 280
                    if v == "total_precipitation_24hr":
                        a = (np.abs(rng.standard_normal((n_t, 1, hh, ww))) * 4e-4).astype(np.float32)   # metres / day
                    arrs[v] = a
                np.savez(os.path.join(d, split, f"2000_{f}.npz"), **arrs)
        np.savez(os.path.join(d, "normalize_mean.npz"), **{v: np.array([280.0 + i]) for i, v in enumerate(vars_)})
        np.savez(os.path.join(d, "normalize_std.npz"), **{v: np.array([3.0 + 0.5 * i]) for i, v in enumerate(vars_)})
        np.save(os.path.join(d, "lat.npy"), np.linspace(60, -60, hh))
        np.save(os.path.join(d, "lon.npy"), np.linspace(0, 350, ww))
    return inp, out


def _ref_module(name):
    from oracle import ref_shim
    root = ref_shim.reference_root()
    spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(root, "src/climate_learn/data", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.reference
@pytest.mark.parametrize("div,overlap", [(1, 4), (2, 2), (2, 3)])
@pytest.mark.parametrize("world", [1, 2])
def test_stream_matches_live_reference(tmp_path, monkeypatch, div, overlap, world):
    from orbit2_b200 import data as D
    import glob
    inp, out = make_shards(str(tmp_path))
    it = _ref_module("iterdataset")
    fi, fo = sorted(glob.glob(inp + "/train/*.npz")), sorted(glob.glob(out + "/train/*.npz"))
    for rank in range(world):
        monkeypatch.setattr(torch.distributed, "get_rank", lambda group=None, r=rank: r)
        monkeypatch.setattr(torch.distributed, "is_initialized", lambda: True)
        ref = it.IndividualDataIter(it.Downscale(it.NpyReader(fi, fo, IN_VARS, OUT_VARS, data_par_size=world, shuffle=False,
                                                              div=div, overlap=overlap)), None, None, subsample=2)
        want = [(torch.stack([x[k] for k in IN_VARS]), torch.stack([y[k] for k in OUT_VARS])) for x, y, _, _ in ref]
        got = list(D.NpzShardStream(fi, fo, IN_VARS, OUT_VARS, rank=rank, world=world, div=div, overlap=overlap, subsample=2))
        assert len(got) == len(want) > 0
        for (gx, gy), (wx, wy) in zip(got, want):
            assert gx.shape == tuple(wx.shape) and gy.shape == tuple(wy.shape)
            assert np.array_equal(gx, wx.numpy()) and np.array_equal(gy, wy.numpy())


@pytest.mark.reference
def test_oracle_transforms_and_metrics_match_live_reference():
    from oracle import ref_shim
    ref = ref_shim.load_reference()
    pm = _ref_module("precipmodule")
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3, 1, 9, 11, generator=g, dtype=torch.float64) * 1e-3
    want = pm.LogTransform(m2mm=True, LOG1P=True, thres_mm_per_day=0.25)(x.clone())
    assert torch.equal(O.log_transform(x), want)
    p = torch.randn(4, 3, 12, 10, generator=g, dtype=torch.float64); t = torch.randn(4, 3, 12, 10, generator=g, dtype=torch.float64)
    lw = O.lat_weights(np.linspace(70, -70, 12))
    fn = ref.functional
    assert torch.allclose(O.rmse(p, t), fn.rmse(p, t), rtol=1e-13)
    assert torch.allclose(O.rmse(p, t, False, lw), fn.rmse(p, t, False, lw), rtol=1e-13)
    assert torch.allclose(O.pearson(p, t), fn.pearson(p, t), rtol=1e-12)
    assert torch.allclose(O.mean_bias(p, t), fn.mean_bias(p, t), rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("n_in,mag,div,overlap", [(32, 4, 2, 4), (48, 2, 4, 3), (16, 4, 1, 4), (64, 1, 4, 2)])
def test_tile_slices_properties(n_in, mag, div, overlap):
    """All tiles have the same size, stay inside the field, cover it, and the output slice is the input slice x mag."""
    from orbit2_b200 import data as D
    top, bottom, _, _ = D.overlap_margins(overlap)
    n_out = n_in * mag
    seen = np.zeros(n_in, bool)
    sizes = set()
    for i in range(div):
        si, so = D.tile_slices(n_in, n_out, div, i, top, bottom)
        assert 0 <= si.start < si.stop <= n_in
        assert (so.start, so.stop) == (si.start * mag, si.stop * mag)
        sizes.add(si.stop - si.start)
        seen[si] = True
    assert len(sizes) == 1 and seen.all()
    if div > 1:
        assert sizes == {n_in // div + top + bottom}


def test_rank_partition_and_wraparound(tmp_path):
    from orbit2_b200 import data as D
    import glob
    inp, out = make_shards(str(tmp_path), n_files=3, n_t=2)
    fi, fo = sorted(glob.glob(inp + "/train/*.npz")), sorted(glob.glob(out + "/train/*.npz"))
    for world in (1, 3, 4):                      # 4 ranks > 3 files: the list wraps around (iterdataset.py:60-66)
        got = [D.NpzShardStream(fi, fo, IN_VARS, OUT_VARS, rank=r, world=world)._files()[0] for r in range(world)]
        assert all(len(g) == max(1, 3 // world) for g in got)
        if world <= 3:
            assert sorted(sum(got, [])) == fi[:len(sum(got, []))]
    # shuffled training lists: every rank must slice the SAME permutation even though each rank has its own sample
    # seed (trainer.npz_data passes seed=rank), so the shards are disjoint and cover every file (iterdataset.py:46-80)
    inp8, out8 = make_shards(str(tmp_path / "eight"), n_files=8, n_t=1)
    f8i, f8o = sorted(glob.glob(inp8 + "/train/*.npz")), sorted(glob.glob(out8 + "/train/*.npz"))
    for world in (2, 4, 8):
        streams = [D.NpzShardStream(f8i, f8o, IN_VARS, OUT_VARS, rank=r, world=world, shuffle=True, seed=r)
                   for r in range(world)]
        orders = []
        for epoch in range(3):
            got = [s._files()[0] for s in streams]
            assert sorted(sum(got, [])) == f8i, (world, epoch, got)
            orders.append(sum(got, []))
        assert orders[0] != orders[1] or orders[1] != orders[2]          # the order changes between epochs
    dm = D.DownscalingData(inp, out, IN_VARS, OUT_VARS, batch_size=2, device="cpu", div=2, overlap=2)
    (_, V, h, w), (_, C, H, W) = dm.get_data_dims()
    x, y = next(iter(D.NpzShardStream(fi, fo, IN_VARS, OUT_VARS, div=2, overlap=2)))
    assert x.shape == (V, h, w) and y.shape == (C, H, W)


@pytest.mark.gpu
def test_device_collate_and_eval_metrics(tmp_path):
    from orbit2_b200 import data as D, losses
    inp, out = make_shards(str(tmp_path), n_files=2, n_t=3, h=10, w=12)
    dm = D.DownscalingData(inp, out, IN_VARS, OUT_VARS, batch_size=4, device="cuda", subsample=1, seed=1)
    mean_i, std_i = dict(np.load(inp + "/normalize_mean.npz")), dict(np.load(inp + "/normalize_std.npz"))
    mean_o, std_o = dict(np.load(out + "/normalize_mean.npz")), dict(np.load(out + "/normalize_std.npz"))
    raw = list(D.NpzShardStream(*dm._lists("val"), IN_VARS, OUT_VARS))
    batches = list(dm.loader("val"))
    assert sum(b[0].shape[0] for b in batches) == len(raw)
    x, y, iv, ov = batches[0]
    assert iv == IN_VARS and ov == OUT_VARS and x.is_cuda
    wx = O.normalize_sample(torch.from_numpy(np.stack([r[0] for r in raw[:x.shape[0]]])), IN_VARS, mean_i, std_i)
    wy = O.normalize_sample(torch.from_numpy(np.stack([r[1] for r in raw[:x.shape[0]]])), OUT_VARS, mean_o, std_o)
    assert torch.allclose(x.cpu(), wx, rtol=1e-6, atol=1e-6) and torch.allclose(y.cpu(), wy, rtol=1e-6, atol=1e-6)
    assert float(y[:, 0].min()) >= 0 and float((y[:, 0] == 0).float().mean()) > 0.05      # dry cells of the log transform
    # evaluation metrics on a noisy "prediction", plain and denormalised, fp32 and bf16 predictions
    g = torch.Generator().manual_seed(3)
    lat = np.load(out + "/lat.npy")
    meta = losses.MetricsMetaInfo(IN_VARS, OUT_VARS, lat, None)
    scale, shift = D.denorm_affine(dm.out_stats)
    for dtype, tol in ((torch.float32, 2e-5), (torch.bfloat16, 2e-2)):
        pred = (y.cpu() + 0.3 * torch.randn(y.shape, generator=g)).to(dtype)
        p64, t64 = pred.double(), y.cpu().double()
        for name, want in (("rmse", O.rmse(p64, t64)), ("lat_rmse", O.rmse(p64, t64, False, O.lat_weights(lat))),
                           ("pearson", O.pearson(p64, t64)), ("mean_bias", O.mean_bias(p64, t64))):
            got = losses.METRICS_REGISTRY[name](aggregate_only=False, metainfo=meta)(pred.cuda(), y)
            assert torch.allclose(got.cpu().double(), want, rtol=tol, atol=tol * 0.05), (name, dtype, got, want)
        sc = torch.from_numpy(scale).double().view(1, -1, 1, 1); sh = torch.from_numpy(shift).double().view(1, -1, 1, 1)
        got = losses.METRICS_REGISTRY["rmse"](aggregate_only=True, metainfo=meta, denorm=(scale, shift))(pred.cuda(), y)
        assert torch.allclose(got.cpu().double(), O.rmse(p64 * sc + sh, t64 * sc + sh, True), rtol=tol)
