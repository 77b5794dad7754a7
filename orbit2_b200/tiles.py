"""TILES spatial tiling: geometry, tiled inference, and the halo exchange for fields sharded across GPUs.

Reference behaviour being mirrored
  * tile extraction    src/climate_learn/data/iterdataset.py:112-177 (every tile has the SAME size
                       (H/div + top + bottom) x (W/div + left + right); edge tiles extend inward)
  * margins            src/climate_learn/data/itermodule.py:161-169: overlap even -> top = bottom = o/2, left = right = 2*(o/2);
                       odd -> top = o//2, bottom = o//2 + 1, left = 2*(o//2), right = 2*(o//2 + 1)
  * stitching          src/climate_learn/utils/visualize.py:125-311: every tile is run through the model on its own and
                       only its inner (H/div*mag) x (W/div*mag) region is copied into the output
                       (the reference has `yo1 -= top ** vmul` at :211, a typo for `*`; the intended form is used here).
The reference runs the div*div tiles sequentially on one rank and has no inter-GPU exchange.  Here the field may be
sharded: rank (v, h) of a (div_v x div_h) grid holds only its inner block, `exchange_halos` fetches the margins from the
ranks that own them (torch.distributed point-to-point: NCCL over NVLink on GPUs, gloo in the CPU tests), each rank runs the
network on its tile, and `gather_output` assembles the inner output blocks.  Attention is per tile (no cross-tile tokens),
exactly like the reference, so the halo exchange of the raw input is the only communication.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def overlap_margins(overlap: int) -> Tuple[int, int, int, int]:
    """(top, bottom, left, right) in low-resolution pixels."""
    if overlap % 2 == 0:
        return overlap // 2, overlap // 2, overlap // 2 * 2, overlap // 2 * 2
    return overlap // 2, overlap // 2 + 1, overlap // 2 * 2, (overlap // 2 + 1) * 2


def axis_bounds(n: int, div: int, idx: int, lo: int, hi: int) -> Tuple[int, int, int, int]:
    """One axis of tile `idx`: input slice [i1, i2) of the field and the inner region [t1, t2) inside the tile.
    lo / hi are the margins before / after (top / bottom or left / right)."""
    if div == 1:
        return 0, n, 0, n
    step = n // div
    i1, i2 = step * idx, step * (idx + 1)
    if idx == 0:
        i2 += lo
    else:
        i1 -= lo
    if idx == div - 1:
        i1 -= hi
    else:
        i2 += hi
    t1 = 0 if idx == 0 else (lo + hi if idx == div - 1 else lo)
    return i1, i2, t1, t1 + step


def tile_size(h: int, w: int, div_v: int, div_h: int, overlap: int) -> Tuple[int, int]:
    top, bottom, left, right = overlap_margins(overlap)
    return (h if div_v == 1 else h // div_v + top + bottom), (w if div_h == 1 else w // div_h + left + right)


def check_tiling(h: int, w: int, div_v: int, div_h: int, overlap: int, patch_size: int):
    """Same admissibility rule as the driver (examples/intermediate_downscaling.py:535-543): the tile must be a
    whole number of patches; additionally the margins must fit inside the neighbouring blocks."""
    th, tw = tile_size(h, w, div_v, div_h, overlap)
    if h % div_v or w % div_h:
        raise ValueError(f"field {h}x{w} is not divisible into {div_v}x{div_h} tiles")
    if th % patch_size or tw % patch_size:
        raise ValueError(f"tile {th}x{tw} is not divisible by patch_size {patch_size}")
    top, bottom, left, right = overlap_margins(overlap)
    if (div_v > 1 and top + bottom > h // div_v) or (div_h > 1 and left + right > w // div_h):
        raise ValueError("overlap margins exceed the tile's own block")
    return th, tw


def tiled_forward(model, x: torch.Tensor, in_variables, out_variables, div: int, overlap: int, mag: int = None,
                  div_h: int = None):
    """Sequential tiled inference on one device (the reference's visualize_at_index loop): returns the stitched
    [B, C, H*mag, W*mag] prediction.  `model.img_size` must already be the tile size (data_config).  The reference tiles
    div x div; ``div_h`` (default: div) allows the div x div_h grids a sharded field uses (2 x 4 over 8 GPUs)."""
    mag = mag or model.superres_mag
    div_v, div_h = div, (div if div_h is None else div_h)
    B, V, H, W = x.shape
    top, bottom, left, right = overlap_margins(overlap)
    out = None
    for v in range(div_v):
        yi1, yi2, yt1, yt2 = axis_bounds(H, div_v, v, top, bottom)
        for h in range(div_h):
            xi1, xi2, xt1, xt2 = axis_bounds(W, div_h, h, left, right)
            pred = model(x[:, :, yi1:yi2, xi1:xi2].contiguous(), in_variables, out_variables)
            if out is None:
                out = torch.empty(B, pred.shape[1], H * mag, W * mag, device=pred.device, dtype=pred.dtype)
            oy, ox = H // div_v * v * mag, W // div_h * h * mag
            out[:, :, oy:oy + (yt2 - yt1) * mag, ox:ox + (xt2 - xt1) * mag] = \
                pred[:, :, yt1 * mag:yt2 * mag, xt1 * mag:xt2 * mag]
    return out


# ------------------------------------------------------------------------------------------------ sharded fields
def _rect_intersect(a, b):
    y1, y2, x1, x2 = max(a[0], b[0]), min(a[1], b[1]), max(a[2], b[2]), min(a[3], b[3])
    return (y1, y2, x1, x2) if y1 < y2 and x1 < x2 else None


class ShardedField:
    """Geometry of a field split into div_v x div_h inner blocks, one per rank (rank = v * div_h + h)."""

    def __init__(self, h: int, w: int, div_v: int, div_h: int, overlap: int):
        self.h, self.w, self.div_v, self.div_h, self.overlap = h, w, div_v, div_h, overlap
        self.top, self.bottom, self.left, self.right = overlap_margins(overlap)
        self.world = div_v * div_h

    def own(self, rank: int):
        v, h = divmod(rank, self.div_h)
        sh, sw = self.h // self.div_v, self.w // self.div_h
        return (sh * v, sh * (v + 1), sw * h, sw * (h + 1))

    def need(self, rank: int):
        v, h = divmod(rank, self.div_h)
        yi1, yi2, _, _ = axis_bounds(self.h, self.div_v, v, self.top, self.bottom)
        xi1, xi2, _, _ = axis_bounds(self.w, self.div_h, h, self.left, self.right)
        return (yi1, yi2, xi1, xi2)

    def inner_in_tile(self, rank: int):
        v, h = divmod(rank, self.div_h)
        _, _, yt1, yt2 = axis_bounds(self.h, self.div_v, v, self.top, self.bottom)
        _, _, xt1, xt2 = axis_bounds(self.w, self.div_h, h, self.left, self.right)
        return (yt1, yt2, xt1, xt2)

    def transfers(self, rank: int):
        """[(peer, rect I receive from peer), ...], [(peer, rect I send to peer), ...] in global coordinates."""
        recv, send = [], []
        for peer in range(self.world):
            if peer == rank:
                continue
            r = _rect_intersect(self.need(rank), self.own(peer))
            if r:
                recv.append((peer, r))
            s = _rect_intersect(self.need(peer), self.own(rank))
            if s:
                send.append((peer, s))
        return recv, send

    def halo_bytes(self, rank: int, channels: int, batch: int = 1, elem: int = 4) -> int:
        recv, _ = self.transfers(rank)
        return sum((r[1] - r[0]) * (r[3] - r[2]) for _, r in recv) * channels * batch * elem


def exchange_halos(x_own: torch.Tensor, geo: ShardedField, rank: int, group=None) -> torch.Tensor:
    """x_own [B, V, H/div_v, W/div_h] (this rank's inner block) -> the rank's full tile [B, V, th, tw] with the margins
    fetched from the owning ranks (one isend / irecv per neighbouring rectangle, batched)."""
    B, V = x_own.shape[:2]
    oy1, oy2, ox1, ox2 = geo.own(rank)
    ny1, ny2, nx1, nx2 = geo.need(rank)
    assert x_own.shape[2] == oy2 - oy1 and x_own.shape[3] == ox2 - ox1, "x_own is not this rank's inner block"
    tile = torch.empty(B, V, ny2 - ny1, nx2 - nx1, device=x_own.device, dtype=x_own.dtype)
    me = _rect_intersect(geo.need(rank), geo.own(rank))
    tile[:, :, me[0] - ny1:me[1] - ny1, me[2] - nx1:me[3] - nx1] = x_own[:, :, me[0] - oy1:me[1] - oy1, me[2] - ox1:me[3] - ox1]
    recv, send = geo.transfers(rank)
    ops, rbufs, keep = [], [], []
    for peer, r in send:
        buf = x_own[:, :, r[0] - oy1:r[1] - oy1, r[2] - ox1:r[3] - ox1].contiguous()
        keep.append(buf)
        ops.append(dist.P2POp(dist.isend, buf, peer if group is None else dist.get_global_rank(group, peer), group))
    for peer, r in recv:
        buf = torch.empty(B, V, r[1] - r[0], r[3] - r[2], device=x_own.device, dtype=x_own.dtype)
        rbufs.append((r, buf))
        ops.append(dist.P2POp(dist.irecv, buf, peer if group is None else dist.get_global_rank(group, peer), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for r, buf in rbufs:
        tile[:, :, r[0] - ny1:r[1] - ny1, r[2] - nx1:r[3] - nx1] = buf
    return tile


def sharded_tiled_forward(model, x_own, in_variables, out_variables, geo: ShardedField, rank: int, group=None):
    """Halo exchange + per-tile forward; returns this rank's inner output block [B, C, H/div_v*mag, W/div_h*mag]."""
    mag = model.superres_mag
    tile = exchange_halos(x_own, geo, rank, group)
    pred = model(tile, in_variables, out_variables)
    yt1, yt2, xt1, xt2 = geo.inner_in_tile(rank)
    return pred[:, :, yt1 * mag:yt2 * mag, xt1 * mag:xt2 * mag].contiguous()


def gather_output(block: torch.Tensor, geo: ShardedField, group=None) -> torch.Tensor:
    """All-gather the inner output blocks into the full [B, C, H*mag, W*mag] field (on every rank)."""
    parts = [torch.empty_like(block) for _ in range(geo.world)]
    dist.all_gather(parts, block, group=group)
    rows = [torch.cat(parts[v * geo.div_h:(v + 1) * geo.div_h], dim=3) for v in range(geo.div_v)]
    return torch.cat(rows, dim=2)
