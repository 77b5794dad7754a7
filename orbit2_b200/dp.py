"""Data-parallel plumbing shared by the training engine and the CPU (gloo) tests: the flat parameter layout and the
bucketed, overlapped gradient all-reduce (reference: FSDP NO_SHARD, examples/intermediate_downscaling.py:618-621 --
gradients averaged over the data-parallel group)."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


class FlatLayout:
    """Offsets of named tensors inside one flat buffer; every slice starts on a multiple of ``align`` elements
    (8 bf16 = 16 bytes: TMA operand bases must be 16-byte aligned)."""

    def __init__(self, names: Sequence[str], sizes: Sequence[int], align: int = 8):
        self.names = list(names)
        self.range: Dict[str, Tuple[int, int]] = {}
        self.padded: Dict[str, Tuple[int, int]] = {}
        total = 0
        for n, s in zip(names, sizes):
            self.range[n] = (total, total + s)
            end = total + (s + align - 1) // align * align
            self.padded[n] = (total, end)
            total = end
        self.total = total

    def runs(self, names: Sequence[str]) -> List[List[int]]:
        """Maximal contiguous [lo, hi) runs covering ``names`` (padding included)."""
        out: List[List[int]] = []
        for lo, hi in sorted(self.padded[n] for n in names):
            if out and out[-1][1] == lo:
                out[-1][1] = hi
            else:
                out.append([lo, hi])
        return out


class BucketReducer:
    """Averages slices of a flat gradient buffer over the process group as soon as they are final.
    ``ready(names)`` enqueues one asynchronous all-reduce per contiguous run; ``finish()`` makes the current stream
    wait for all of them.  With world size 1 both are no-ops."""

    def __init__(self, flat_grad: torch.Tensor, layout: FlatLayout, group=None):
        self.flat, self.layout, self.group = flat_grad, layout, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.pending = []
        self.reduced_elems = 0
        # gloo has no AVG: sum then scale
        self._avg = dist.ReduceOp.AVG if (self.world > 1 and dist.get_backend(group) == "nccl") else None

    def ready(self, names: Sequence[str]):
        if self.world == 1:
            return
        for lo, hi in self.layout.runs(names):
            t = self.flat[lo:hi]
            if self._avg is not None:
                self.pending.append((dist.all_reduce(t, op=self._avg, group=self.group, async_op=True), None))
            else:
                self.pending.append((dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True), t))
            self.reduced_elems += hi - lo

    def finish(self):
        for work, t in self.pending:
            work.wait()
            if t is not None:
                t.div_(self.world)
        self.pending = []


class ShardedReducer:
    """FSDP-style gradient reduction for a flat buffer whose optimizer state is sharded evenly over the ranks (rank r owns
    elements [r*S, (r+1)*S)): every finished slice is reduced (averaged) onto the rank(s) that own it -- a reduce-scatter
    issued bucket by bucket -- so that no rank ever holds, or updates, more than its own shard of the summed gradient
    (reference: FSDP FULL_SHARD, examples/intermediate_downscaling.py:615-617)."""

    def __init__(self, flat_grad: torch.Tensor, layout: FlatLayout, group=None):
        self.flat, self.layout, self.group = flat_grad, layout, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.shard = shard_size(layout.total, self.world)
        self.own = (self.rank * self.shard, (self.rank + 1) * self.shard)
        self.pending = []
        self.reduced_elems = 0
        self._nccl = self.world > 1 and dist.get_backend(group) == "nccl"

    def segments(self, lo: int, hi: int):
        """[(owner, seg_lo, seg_hi)] of a run split at shard boundaries."""
        out = []
        r = lo // self.shard
        while lo < hi:
            end = min(hi, (r + 1) * self.shard)
            out.append((r, lo, end))
            lo, r = end, r + 1
        return out

    def ready(self, names: Sequence[str]):
        if self.world == 1:
            return
        for lo, hi in self.layout.runs(names):
            for owner, a, b in self.segments(lo, hi):
                t = self.flat[a:b]
                dst = owner if self.group is None else dist.get_global_rank(self.group, owner)
                op = dist.ReduceOp.AVG if self._nccl else dist.ReduceOp.SUM
                self.pending.append((dist.reduce(t, dst=dst, op=op, group=self.group, async_op=True),
                                     t if (not self._nccl and owner == self.rank) else None))
                self.reduced_elems += b - a

    def finish(self):
        for work, t in self.pending:
            work.wait()
            if t is not None:
                t.div_(self.world)
        self.pending = []


def shard_size(total: int, world: int, align: int = 8) -> int:
    per = (total + world - 1) // world
    return (per + align - 1) // align * align
