"""CPU, world_size 2, gloo: the bucketed gradient all-reduce used by the training engine averages every slice of the
flat gradient buffer exactly once, whatever order / grouping the backward schedule reports them in."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from orbit2_b200.dp import BucketReducer, FlatLayout
    from orbit2_b200.reslim import kernel_param_names
    names = ["var_embed", "var_query"] + kernel_param_names(depth=2, dec=1) + ["token_embeds.0.proj.weight"]
    g = torch.Generator().manual_seed(0)
    sizes = [int(torch.randint(1, 50, (1,), generator=g)) for _ in names]
    lay = FlatLayout(names, sizes)
    flat = torch.arange(lay.total, dtype=torch.float32) * (rank + 1)          # rank-dependent "gradients"
    red = BucketReducer(flat, lay)
    # the order reslim_backward reports groups in: convs, head, final norm, blocks (reversed), var_agg.proj, then the rest
    kn = kernel_param_names(2, 1)
    groups = [[n for n in kn if n.startswith(("conv_out", "path2"))], [n for n in kn if n.startswith("head.")],
              ["norm.weight", "norm.bias"], [n for n in kn if n.startswith("blocks.1.")],
              [n for n in kn if n.startswith("blocks.0.")], ["var_agg.proj.weight", "var_agg.proj.bias"],
              ["var_embed", "var_query", "token_embeds.0.proj.weight"]]
    for grp in groups:
        red.ready(grp)
    red.finish()
    expect = torch.arange(lay.total, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(flat, expect) and red.reduced_elems == lay.total
    # FSDP-style: every slice is reduced onto its owner only; each rank ends with the averaged gradient of ITS shard
    from orbit2_b200.dp import ShardedReducer, shard_size
    S = shard_size(lay.total, world)
    flat2 = torch.zeros(S * world)
    flat2[:lay.total] = torch.arange(lay.total, dtype=torch.float32) * (rank + 1)
    sred = ShardedReducer(flat2, lay)
    assert sred.shard == S and sred.own == (rank * S, (rank + 1) * S)
    for grp in groups:
        sred.ready(grp)
    sred.finish()
    o0, o1 = sred.own
    o1 = min(o1, lay.total)
    ok = ok and torch.allclose(flat2[o0:o1], expect[o0:o1]) and sred.reduced_elems == lay.total
    segs = sred.segments(S - 3, S + 5)
    ok = ok and segs == [(0, S - 3, S), (1, S, S + 5)]
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_bucket_reducer_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


# ---------------------------------------------------------------------------------------------------------------------
# FULL_SHARD (dp.FullShard): per-unit all-gather / reduce-scatter driven by the lookups of the kernel schedule
def _fs_units(depth, dec, D):
    g = torch.Generator().manual_seed(5)
    shapes = {"var_agg.proj.weight": (D, D)}
    for j in range(dec + 1):
        shapes[f"head.{2 * j}.weight"] = (D, D) if j < dec else (3 * D + 1, D)        # ragged sizes exercise the padding
    for i in range(depth):
        b = f"blocks.{i}."
        shapes.update({b + "attn.qkv.weight": (3 * D, D), b + "attn.proj.weight": (D, D), b + "mlp.fc1.weight": (4 * D, D),
                       b + "mlp.fc2.weight": (D, 4 * D)})
    tensors = {n: torch.randn(*s, generator=g) for n, s in shapes.items()}
    root = [n for n in shapes if not n.startswith("blocks.")]
    units = [root] + [[n for n in shapes if n.startswith(f"blocks.{i}.")] for i in range(depth)]
    return units, tensors


def _fs_schedule(fs, units, tensors, rank, world, dtype, ckpt):
    """Walk the lookups reslim_forward / reslim_backward make; returns True when every gathered weight equals the full
    tensor and every gradient shard equals the rank-average of the full gradient."""
    depth = len(units) - 1
    cast = (lambda t: t.to(dtype))
    ok = True
    fs.begin(+1)
    ok &= torch.equal(fs.params["var_agg.proj.weight"], cast(tensors["var_agg.proj.weight"]))
    for i in range(1, depth + 1):
        for n in units[i]:
            ok &= torch.equal(fs.params[n], cast(tensors[n]))
    for n in units[0]:
        ok &= torch.equal(fs.params[n], cast(tensors[n]))
    fs.begin(-1)
    gfull = {n: torch.randn(tensors[n].shape, generator=torch.Generator().manual_seed(len(n) + 17)) for n in tensors}
    heads = [n for n in units[0] if n.startswith("head.")]
    for n in heads:
        fs.grads[n].add_(gfull[n] * (rank + 1))
        ok &= torch.equal(fs.params[n], cast(tensors[n]))
    fs.ready(heads)
    for i in range(depth, 0, -1):
        if ckpt:                                    # the recomputed forward reads the unit's weights again
            for n in units[i]:
                ok &= torch.equal(fs.params[n], cast(tensors[n]))
        for n in reversed(units[i]):
            fs.grads[n].add_(gfull[n] * (rank + 1))      # accumulate twice, like split-K partial tiles do
            fs.grads[n].add_(gfull[n] * (rank + 1))
            ok &= torch.equal(fs.params[n], cast(tensors[n]))
        fs.ready(units[i] + [f"blocks.{i - 1}.norm1.bias"])          # names the shard does not own are ignored
    fs.grads["var_agg.proj.weight"].add_(gfull["var_agg.proj.weight"] * (rank + 1))
    fs.ready(["var_agg.proj.weight", "var_agg.proj.bias"])
    fs.finish()
    avg = sum(range(1, world + 1)) / world
    for u, names in enumerate(units):
        S = fs.S[u]
        full = torch.zeros(S * world)
        for n in names:
            lo = fs.where[n][1]
            full[lo:lo + gfull[n].numel()] = gfull[n].reshape(-1) * avg * (1 if u == 0 else 2)
        ok &= torch.allclose(fs.gshard[u], full[rank * S:(rank + 1) * S], rtol=1e-6, atol=1e-6)
    return bool(ok)


def _fs_worker(rank, world, port, ret):
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from orbit2_b200.dp import FullShard
    ok = True
    for dtype, ckpt in ((torch.float32, False), (torch.bfloat16, True)):
        units, tensors = _fs_units(depth=5, dec=1, D=12)
        fs = FullShard(units, tensors, dtype)
        assert fs.world == world and all(s % 8 == 0 for s in fs.S)
        total = sum(t.numel() for t in tensors.values())
        assert total <= sum(fs.S) * world < total + (8 * world + 8) * len(units) + 8 * len(tensors)
        for step in range(2):
            ok &= _fs_schedule(fs, units, tensors, rank, world, dtype, ckpt)
            # each rank updates ITS shard only (the fused AdamW stands in as: p -= 0.5 * g, low-precision copy refreshed)
            for p32, g32, m, v, low in fs.shards():
                p32.sub_(0.5 * g32)
                if low is not None:
                    low.copy_(p32.to(low.dtype))
            fs.after_step()
            avg = sum(range(1, world + 1)) / world
            for u, names in enumerate(units):
                for n in names:
                    gf = torch.randn(tensors[n].shape, generator=torch.Generator().manual_seed(len(n) + 17))
                    tensors[n] = tensors[n] - 0.5 * gf * avg * (1 if u == 0 else 2)
            ok &= all(torch.allclose(fs.full_tensor(n), tensors[n], rtol=1e-6, atol=1e-6) for n in ("head.2.weight",
                                                                                                  "blocks.3.mlp.fc2.weight"))
            tensors = {n: fs.full_tensor(n) for n in tensors}          # exact fp32 masters for the next round's equality
        # gathers: root at construction and after every update; per step every Block once per direction (the last two Blocks stay in their slots at the
        # forward -> backward turn; a checkpointed Block's recomputation reuses the gathered unit)
        blk = sum(fs.S[1:]) * world
        assert fs.gathered_elems == 3 * fs.S[0] * world + 2 * (2 * blk - (fs.S[5] + fs.S[4]) * world), fs.gathered_elems
        assert fs.scattered_elems == 2 * sum(fs.S) * world
    ret[rank] = bool(ok)
    if world > 1:
        dist.destroy_process_group()


def test_full_shard_world1():
    ret = {}
    _fs_worker(0, 1, 0, ret)
    assert ret == {0: True}


def test_full_shard_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fs_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
