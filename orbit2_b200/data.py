"""Data path either side of the kernels: the reference's npz shard format read on the host, the
per-variable normalisation done on the GPU.

Reference pipeline (src/climate_learn/data): ``NpyReader`` (iterdataset.py:21-177: files split over ranks, optional TILES
extraction with halos) -> ``Downscale`` (:318-330, to float32) -> ``IndividualDataIter`` (:333-383: one sample every
``subsample`` time steps, per-variable ``Normalize`` / ``LogTransform``) -> ``ShuffleIterableDataset`` (:386-405) ->
``DataLoader(collate_fn)`` (itermodule.py:451-469: variables stacked to [B,V,H,W]).  On-disk layout (itermodule.py:83-126,
202-231): ``{root}/{train,val,test}/*.npz`` with one ``[N,1,H,W]`` array per variable, ``{root}/normalize_{mean,std}.npz``,
``{root}/lat.npy`` / ``lon.npy``, ``{root}/{split}/climatology.npz``.

Here the host only slices and stacks RAW fields into pinned staging buffers; the copy to the device carries raw data and one
kernel (``o2_normalize_fields``) applies every variable's transform in place.  One process per GPU (torchrun): the rank's
share of the files is ``rank * per_rank .. (rank + 1) * per_rank`` exactly like the reference with one loader worker.
"""
from __future__ import annotations

import glob
import os
import random
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

# src/climate_learn/data/processing/era5_constants.py (PRECIP_VARIABLES)
PRECIP_VARIABLES = ("total_precipitation_24hr", "total_precipitation")


def tile_slices(n_in: int, n_out: int, div: int, index: int, before: int, after: int) -> Tuple[slice, slice]:
    """1-D TILES slice of tile ``index`` (iterdataset.py:112-170): every tile has ``n_in // div + before + after`` input
    cells; the first tile extends only inward, the last only backward, so all tiles have the same size."""
    if div == 1:
        return slice(0, n_in), slice(0, n_out)
    mul = n_out // n_in
    i1, i2 = n_in // div * index, n_in // div * (index + 1)
    o1, o2 = n_out // div * index, n_out // div * (index + 1)
    if index == 0:
        i2 += before; o2 += before * mul
    else:
        i1 -= before; o1 -= before * mul
    if index == div - 1:
        i1 -= after; o1 -= after * mul
    else:
        i2 += after; o2 += after * mul
    return slice(i1, i2), slice(o1, o2)


def overlap_margins(overlap: int) -> Tuple[int, int, int, int]:
    """(top, bottom, left, right) halo of the TILES split in low-resolution cells (iterdataset.py:100-110)."""
    if overlap % 2 == 0:
        return overlap // 2, overlap // 2, overlap // 2 * 2, overlap // 2 * 2
    return overlap // 2, overlap // 2 + 1, overlap // 2 * 2, (overlap // 2 + 1) * 2


class NpzShardStream:
    """Raw (un-normalised) samples ``(x [V,h,w], y [C,H,W])`` of this rank's shard files, in the reference's order."""

    def __init__(self, inp_files: Sequence[str], out_files: Sequence[str], in_vars: Sequence[str], out_vars: Sequence[str],
                 rank: int = 0, world: int = 1, div: int = 1, overlap: int = 4, subsample: int = 1, shuffle: bool = False,
                 buffer_size: int = 0, seed: Optional[int] = None, file_seed: int = 0):
        assert len(inp_files) == len(out_files)
        self.inp_files = [f for f in inp_files if "climatology" not in f]
        self.out_files = [f for f in out_files if "climatology" not in f]
        self.in_vars, self.out_vars = list(in_vars), list(out_vars)
        self.rank, self.world, self.div, self.overlap = rank, world, div, overlap
        self.subsample, self.shuffle, self.buffer_size = subsample, shuffle, buffer_size
        # Two generators.  The FILE order must be the same permutation on every rank, otherwise the per-rank slices
        # below overlap and some files are never read (the reference shuffles after seed_everything(0) on all ranks,
        # iterdataset.py:46-80); it is seeded by ``file_seed`` alone and advances once per pass, so every epoch draws a
        # new order that all ranks agree on.  The shuffle BUFFER is private to the rank (``seed``).
        self.rng = random.Random(seed)
        self.file_rng = random.Random(file_seed)

    def _files(self):
        inp, out = list(self.inp_files), list(self.out_files)
        if self.shuffle:
            order = list(range(len(inp)))
            self.file_rng.shuffle(order)
            inp, out = [inp[i] for i in order], [out[i] for i in order]
        n = len(inp)
        if n < self.world:                          # wrap the list around when there are fewer files than ranks
            mult, rem = self.world // n, self.world - n * (self.world // n)
            inp, out = inp * mult + inp[:rem], out * mult + out[:rem]
            n = len(inp)
        per = n // self.world
        return inp[self.rank * per:(self.rank + 1) * per], out[self.rank * per:(self.rank + 1) * per]

    def _tiles(self):
        inp_files, out_files = self._files()
        top, bottom, left, right = overlap_margins(self.overlap)
        for pi, po in zip(inp_files, out_files):
            inp = np.load(pi)
            out = inp if po == pi else np.load(po)
            xs = {k: inp[k] for k in self.in_vars}                      # [N,1,h,w]
            ys = {k: out[k] for k in self.out_vars}
            h, w = xs[self.in_vars[0]].shape[2:]
            H, W = ys[self.out_vars[0]].shape[2:]
            for vi in range(self.div):
                yi, yo = tile_slices(h, H, self.div, vi, top, bottom)
                for hi in range(self.div):
                    xi, xo = tile_slices(w, W, self.div, hi, left, right)
                    yield ({k: v[:, 0, yi, xi] for k, v in xs.items()}, {k: v[:, 0, yo, xo] for k, v in ys.items()})

    def _samples(self):
        for xs, ys in self._tiles():
            n = xs[self.in_vars[0]].shape[0]
            assert all(v.shape[0] == n for v in xs.values()) and all(v.shape[0] == n for v in ys.values())
            for i in range(0, n, self.subsample):
                yield (np.stack([xs[k][i] for k in self.in_vars]).astype(np.float32, copy=False),
                       np.stack([ys[k][i] for k in self.out_vars]).astype(np.float32, copy=False))

    def __iter__(self) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
        if self.buffer_size <= 0:
            yield from self._samples()
            return
        buf: List = []                              # ShuffleIterableDataset, iterdataset.py:386-405
        for s in self._samples():
            if len(buf) == self.buffer_size:
                idx = self.rng.randint(0, self.buffer_size - 1)
                yield buf[idx]
                buf[idx] = s
            else:
                buf.append(s)
        self.rng.shuffle(buf)
        while buf:
            yield buf.pop()


class DeviceCollator:
    """Stacks ``batch_size`` raw samples into pinned staging buffers, copies them to the GPU asynchronously and normalises
    them there (one launch per tensor).  Yields ``(x [B,V,h,w], y [B,C,H,W], in_vars, out_vars)`` like the reference's
    collate (itermodule.py:451-469)."""

    def __init__(self, stream, batch_size: int, in_stats, out_stats, device, drop_last: bool = False):
        self.stream, self.batch_size, self.device, self.drop_last = stream, batch_size, torch.device(device), drop_last
        self.in_stats = tuple(t.to(self.device) for t in in_stats)       # (mean [V], std [V], kind [V])
        self.out_stats = tuple(t.to(self.device) for t in out_stats)
        self._pin = None

    def _emit(self, xs, ys):
        B = len(xs)
        if self._pin is None or self._pin[0].shape[1:] != (len(xs[0]),) + xs[0].shape[1:] or self._pin[0].shape[0] < B:
            self._pin = (torch.empty((self.batch_size,) + xs[0].shape, dtype=torch.float32).pin_memory(),
                         torch.empty((self.batch_size,) + ys[0].shape, dtype=torch.float32).pin_memory())
        px, py = self._pin
        for i in range(B):
            px[i].copy_(torch.from_numpy(xs[i]))
            py[i].copy_(torch.from_numpy(ys[i]))
        x = px[:B].to(self.device, non_blocking=True)
        y = py[:B].to(self.device, non_blocking=True)
        ops.normalize_fields_(x, *self.in_stats)
        ops.normalize_fields_(y, *self.out_stats)
        torch.cuda.current_stream().synchronize()       # the staging buffers are reused by the next batch
        return x, y, list(self.stream.in_vars), list(self.stream.out_vars)

    def __iter__(self):
        xs, ys = [], []
        for x, y in self.stream:
            xs.append(x); ys.append(y)
            if len(xs) == self.batch_size:
                yield self._emit(xs, ys)
                xs, ys = [], []
        if xs and not self.drop_last:
            yield self._emit(xs, ys)


def normalize_stats(root_dir: str, variables: Sequence[str]):
    """(mean [V], std [V], kind [V]) from ``normalize_{mean,std}.npz`` (itermodule.py:202-211): precipitation variables use
    the log transform (kind 1) instead of mean / std."""
    mean = dict(np.load(os.path.join(root_dir, "normalize_mean.npz")))
    std = dict(np.load(os.path.join(root_dir, "normalize_std.npz")))
    m, s, k = [], [], []
    for v in variables:
        if v in PRECIP_VARIABLES:
            m.append(0.0); s.append(1.0); k.append(1)
        else:
            m.append(float(mean[v][0])); s.append(float(std[v][0])); k.append(0)
    return (torch.tensor(m, dtype=torch.float32), torch.tensor(s, dtype=torch.float32), torch.tensor(k, dtype=torch.int32))


def denorm_affine(out_stats):
    """(scale [C], shift [C]) of the reference's denormalising transform for the evaluation metrics
    (Normalize(-mean/std, 1/std), i.e. x * std + mean); log-transformed channels stay in log space, scale 1 / shift 0."""
    mean, std, kind = out_stats
    scale = torch.where(kind == 0, std, torch.ones_like(std))
    shift = torch.where(kind == 0, mean, torch.zeros_like(mean))
    return scale.numpy(), shift.numpy()


class DownscalingData:
    """The slice of IterDataModule (itermodule.py:42-240) the downscaling driver uses."""

    def __init__(self, inp_root_dir: str, out_root_dir: str, in_vars: Sequence[str], out_vars: Sequence[str],
                 batch_size: int, device, rank: int = 0, world: int = 1, div: int = 1, overlap: int = 4, subsample: int = 1,
                 buffer_size: int = 0, seed: Optional[int] = 0, file_seed: int = 0):
        self.file_seed, self._epoch = file_seed, {}
        self.inp_root_dir, self.out_root_dir = inp_root_dir, out_root_dir
        self.in_vars, self.out_vars = list(in_vars), list(out_vars)
        self.batch_size, self.device, self.rank, self.world = batch_size, device, rank, world
        self.div, self.overlap, self.subsample, self.buffer_size, self.seed = div, overlap, subsample, buffer_size, seed
        self.in_stats = normalize_stats(inp_root_dir, self.in_vars)
        self.out_stats = normalize_stats(out_root_dir, self.out_vars)

    def _lists(self, split):
        return (sorted(glob.glob(os.path.join(self.inp_root_dir, split, "*.npz"))),
                sorted(glob.glob(os.path.join(self.out_root_dir, split, "*.npz"))))

    def get_lat_lon(self):
        return (np.load(os.path.join(self.out_root_dir, "lat.npy")), np.load(os.path.join(self.out_root_dir, "lon.npy")))

    def get_data_dims(self):
        """((B, V, h, w), (B, C, H, W)) of one (tile of a) sample, itermodule.py:137-199."""
        h = len(np.load(os.path.join(self.inp_root_dir, "lat.npy"))); w = len(np.load(os.path.join(self.inp_root_dir, "lon.npy")))
        H = len(np.load(os.path.join(self.out_root_dir, "lat.npy"))); W = len(np.load(os.path.join(self.out_root_dir, "lon.npy")))
        if self.div > 1:
            top, bottom, left, right = overlap_margins(self.overlap)
            h2, w2 = h // self.div + top + bottom, w // self.div + left + right
            H, W = H // self.div + (top + bottom) * (H // h), W // self.div + (left + right) * (W // w)
            h, w = h2, w2
        return (self.batch_size, len(self.in_vars), h, w), (self.batch_size, len(self.out_vars), H, W)

    def get_climatology(self, split="val"):
        clim = np.load(os.path.join(self.out_root_dir, split, "climatology.npz"))
        return {v: torch.from_numpy(np.squeeze(clim[v].astype(np.float32), axis=0)) for v in self.out_vars}

    def loader(self, split: str, shuffle: Optional[bool] = None):
        train = split == "train"
        inp, out = self._lists(split)
        stream = NpzShardStream(inp, out, self.in_vars, self.out_vars, self.rank, self.world, self.div, self.overlap,
                                self.subsample, shuffle=train if shuffle is None else shuffle,
                                buffer_size=self.buffer_size if train else 0, seed=self.seed,
                                file_seed=self.file_seed + 7919 * self._epoch.get(split, 0))
        self._epoch[split] = self._epoch.get(split, 0) + 1            # same count on every rank -> same file order
        return DeviceCollator(stream, self.batch_size, self.in_stats, self.out_stats, self.device)
