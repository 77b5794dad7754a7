"""Data-parallel plumbing shared by the training engine and the CPU (gloo) tests: the flat parameter layout and the
bucketed, overlapped gradient all-reduce (reference: FSDP NO_SHARD, examples/intermediate_downscaling.py:618-621 --
gradients averaged over the data-parallel group)."""
from __future__ import annotations

from typing import Optional, Dict, List, Sequence, Tuple

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


class FullShard:
    """FSDP FULL_SHARD for the GEMM weights (reference: examples/intermediate_downscaling.py:583-617 -- every Block is one
    FSDP unit, its flat parameter is all-gathered before the unit's forward and again before its backward, its gradient
    is reduce-scattered when the unit's backward ends; fp32 master copy, Adam state and the update live on the shard).

    ``units[0]`` is the root unit (var_agg.proj + head weights: needed first in forward, last in backward -> kept gathered
    for the whole step); ``units[1:]`` are the Blocks in execution order.  Per rank and unit the persistent state is one
    1/world slice of the unit's flat buffer: fp32 master, fp32 gradient, Adam m / v and (bf16 arm) the bf16 copy the fused
    AdamW kernel refreshes.  Block weights are gathered in the compute dtype into two rotating slots: the first access of
    unit i waits for its gather and at once enqueues the gather of the next unit in the current direction (i + 1 in
    forward, i - 1 in backward, the recomputed forward of a checkpointed Block included), so the exchange runs under the
    unit's kernels.  Full-size gradients exist for two units at a time (fp32 slots); ``ready()`` reduce-scatters a slot
    (average) into the gradient shard as soon as all the unit's weights are reported final.

    ``params`` / ``grads`` are read-only mappings name -> tensor view; looking a name up is what triggers the gather /
    the slot hand-over, so the kernel schedule (reslim_forward / reslim_backward) needs no knowledge of the sharding.
    Works on any device and backend (gloo on CPU in the tests, NCCL on the GPUs); world size 1 copies instead of
    gathering."""

    def __init__(self, units: Sequence[Sequence[str]], tensors: Dict[str, torch.Tensor], compute_dtype: torch.dtype,
                 group=None):
        self.group = group
        multi = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._nccl = self.world > 1 and dist.get_backend(group) == "nccl"
        self.dtype = compute_dtype
        self.lowp = compute_dtype != torch.float32
        self.units = [list(u) for u in units]
        dev = next(iter(tensors.values())).device
        self.device = dev
        self.where: Dict[str, Tuple[int, int, torch.Size]] = {}
        self.layouts, self.S = [], []
        self.master, self.gshard, self.m, self.v, self.low = [], [], [], [], []
        for u, names in enumerate(self.units):
            lay = FlatLayout(names, [tensors[n].numel() for n in names], align=8)
            S = shard_size(lay.total, self.world)
            self.layouts.append(lay)
            self.S.append(S)
            full = torch.zeros(S * self.world, device=dev, dtype=torch.float32)
            for n in names:
                lo, hi = lay.range[n]
                full[lo:hi].copy_(tensors[n].detach().reshape(-1))
                self.where[n] = (u, lo, tensors[n].shape)
            own = full[self.rank * S:(self.rank + 1) * S].clone()
            del full
            self.master.append(own)
            self.gshard.append(torch.zeros_like(own))
            self.m.append(torch.zeros_like(own))
            self.v.append(torch.zeros_like(own))
            self.low.append(own.to(compute_dtype) if self.lowp else None)
        n_root = self.S[0] * self.world
        n_blk = max([s * self.world for s in self.S[1:]], default=0)
        self.root_w = torch.zeros(n_root, device=dev, dtype=compute_dtype)
        self.root_g = torch.zeros(n_root, device=dev, dtype=torch.float32)
        self.w_slots = [torch.zeros(n_blk, device=dev, dtype=compute_dtype) for _ in range(2 if n_blk else 0)]
        self.g_slots = [torch.zeros(n_blk, device=dev, dtype=torch.float32) for _ in range(2 if n_blk else 0)]
        self.w_owner = [None, None]             # unit whose gathered weights a slot holds (or is receiving)
        self.root_valid = False
        self.w_work: Dict[int, object] = {}     # unit -> outstanding gather
        self.g_owner = [None, None]             # unit whose gradient a slot is accumulating
        self.g_work = [None, None]              # outstanding reduce-scatter out of a slot
        self.g_live: Dict[int, torch.Tensor] = {}
        self.g_ready: Dict[int, set] = {}
        self.pending = []
        self.direction = +1
        self.gathered_elems = 0
        self.scattered_elems = 0
        self.params = _LazyMap(self.param, self.where)
        self.grads = _LazyMap(self.grad, self.where)
        self._gather(0)

    # ---- parameters
    def _w_buf(self, u: int) -> torch.Tensor:
        return self.root_w if u == 0 else self.w_slots[u % 2]

    def _gather(self, u: int):
        """Enqueue the all-gather of unit u's weights (no-op when its buffer already holds, or is receiving, them)."""
        if u == 0:
            if self.root_valid:
                return
            self.root_valid = True
        elif self.w_owner[u % 2] == u:
            return
        else:
            self.w_owner[u % 2] = u
        src = self.low[u] if self.lowp else self.master[u]
        dst = self._w_buf(u)[:self.S[u] * self.world]
        work = None
        if self.world > 1:
            work = dist.all_gather_into_tensor(dst, src, group=self.group, async_op=True)
        else:
            dst.copy_(src)
        self.gathered_elems += dst.numel()
        self.w_work[u] = work

    def begin(self, direction: int):
        """+1 before the forward schedule, -1 before the backward schedule (sets the prefetch direction)."""
        self.direction = direction
        if direction > 0 and len(self.units) > 1:
            self._gather(1)

    def param(self, name: str) -> torch.Tensor:
        u, lo, shape = self.where[name]
        self._gather(u)
        work = self.w_work.pop(u, None)
        if work is not None:
            work.wait()
        if u != 0:                                 # prefetch the neighbour the schedule visits next
            nxt = u + self.direction
            if 1 <= nxt < len(self.units) and self.w_owner[nxt % 2] != nxt:
                self._gather(nxt)
        return self._w_buf(u)[lo:lo + shape.numel()].view(shape)

    # ---- gradients
    def grad(self, name: str) -> torch.Tensor:
        u, lo, shape = self.where[name]
        buf = self.g_live.get(u)
        if buf is None:
            if u == 0:
                buf = self.root_g
            else:
                s = u % 2
                if self.g_work[s] is not None:     # the slot's previous reduce-scatter must have drained it
                    self.g_work[s].wait()
                    self.g_work[s] = None
                self.g_owner[s] = u
                buf = self.g_slots[s]
            buf = buf[:self.S[u] * self.world]
            buf.zero_()
            self.g_live[u] = buf
            self.g_ready[u] = set()
        return buf[lo:lo + shape.numel()].view(shape)

    def ready(self, names: Sequence[str]):
        """Report gradients final; a unit whose weights are all final is reduce-scattered (averaged) into its shard."""
        for n in names:
            if n not in self.where:
                continue
            u = self.where[n][0]
            if u not in self.g_live:
                self.grad(n)                       # a unit that produced no gradient still contributes zeros
            self.g_ready[u].add(n)
            if len(self.g_ready[u]) == len(self.units[u]):
                self._scatter(u)

    def _scatter(self, u: int):
        buf = self.g_live.pop(u)
        self.g_ready.pop(u)
        out = self.gshard[u]
        if self.world > 1:
            op = dist.ReduceOp.AVG if self._nccl else dist.ReduceOp.SUM
            work = dist.reduce_scatter_tensor(out, buf, op=op, group=self.group, async_op=True)
            self.pending.append((work, None if self._nccl else out))
            if u != 0:
                self.g_work[u % 2] = work
        else:
            out.copy_(buf)
        self.scattered_elems += buf.numel()

    def finish(self):
        """Make the current stream wait for every outstanding reduce-scatter (call before the optimizer)."""
        assert not self.g_live, f"units {sorted(self.g_live)} were never reported ready"
        for work, t in self.pending:
            work.wait()
            if t is not None:
                t.div_(self.world)
        self.pending = []
        self.g_work = [None, None]

    # ---- optimizer side
    def shards(self):
        """(fp32 master, fp32 gradient, m, v, low-precision copy or None) per unit -- what the fused AdamW updates."""
        return list(zip(self.master, self.gshard, self.m, self.v, self.low))

    def after_step(self):
        """The shards changed: every gathered copy is stale; the root unit (and the first Block) are re-gathered now."""
        for work in self.w_work.values():
            if work is not None:
                work.wait()
        self.w_work.clear()
        self.w_owner = [None, None]
        self.root_valid = False
        self._gather(0)

    def full_tensor(self, name: str) -> torch.Tensor:
        """fp32 master value of one weight, gathered from all ranks (checkpoints, tests); collective."""
        u, lo, shape = self.where[name]
        full = torch.empty(self.S[u] * self.world, device=self.device, dtype=torch.float32)
        if self.world > 1:
            dist.all_gather_into_tensor(full, self.master[u], group=self.group)
        else:
            full.copy_(self.master[u])
        return full[lo:lo + shape.numel()].view(shape).clone()

    def _gather_unit(self, shard: torch.Tensor, u: int) -> torch.Tensor:
        full = torch.empty(self.S[u] * self.world, device=self.device, dtype=shard.dtype)
        if self.world > 1:
            dist.all_gather_into_tensor(full, shard, group=self.group)
        else:
            full.copy_(shard)
        return full

    def full_adam_state(self) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
        """name -> (exp_avg, exp_avg_sq), fp32 on the host, gathered unit by unit (checkpoints); collective."""
        out = {}
        for u, names in enumerate(self.units):
            m, v = self._gather_unit(self.m[u], u).cpu(), self._gather_unit(self.v[u], u).cpu()
            for n in names:
                _, lo, shape = self.where[n]
                out[n] = (m[lo:lo + shape.numel()].view(shape).clone(), v[lo:lo + shape.numel()].view(shape).clone())
        return out

    def load_tensor(self, name: str, value: Optional[torch.Tensor] = None, exp_avg: Optional[torch.Tensor] = None,
                    exp_avg_sq: Optional[torch.Tensor] = None):
        """Overwrite this rank's slice of a weight (and / or its Adam moments) from FULL tensors (checkpoint resume);
        local, every rank calls it with the same values.  ``after_step()`` must follow to re-gather the operands."""
        u, lo, shape = self.where[name]
        S, r = self.S[u], self.rank
        a, b = max(lo, r * S), min(lo + shape.numel(), (r + 1) * S)      # overlap of the tensor with this rank's slice
        if a >= b:
            return
        for src, dst in ((value, self.master[u]), (exp_avg, self.m[u]), (exp_avg_sq, self.v[u])):
            if src is not None:
                dst[a - r * S:b - r * S].copy_(src.reshape(-1)[a - lo:b - lo].to(dst.device, torch.float32))
        if value is not None and self.lowp:
            self.low[u][a - r * S:b - r * S].copy_(self.master[u][a - r * S:b - r * S])

    def persistent_bytes(self) -> int:
        per = sum(self.S) * (16 + (2 if self.lowp else 0))
        slots = sum(t.numel() * t.element_size() for t in [self.root_w, self.root_g] + self.w_slots + self.g_slots)
        return per + slots


class _LazyMap:
    """Read-only name -> tensor mapping whose lookups call ``getter`` (FullShard.param / FullShard.grad)."""

    def __init__(self, getter, names):
        self._get, self._names = getter, names

    def __getitem__(self, name):
        if name not in self._names:
            raise KeyError(name)
        return self._get(name)

    def __contains__(self, name):
        return name in self._names

    def keys(self):
        return self._names.keys()

    def __iter__(self):
        return iter(self._names)

    def __len__(self):
        return len(self._names)


class ChainedMap:
    """Lookup in the first mapping that knows the name (replicated parameters first, sharded ones second)."""

    def __init__(self, *maps):
        self.maps = maps

    def __getitem__(self, name):
        for m in self.maps:
            if name in m:
                return m[name]
        raise KeyError(name)

    def __contains__(self, name):
        return any(name in m for m in self.maps)
