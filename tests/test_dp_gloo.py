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
