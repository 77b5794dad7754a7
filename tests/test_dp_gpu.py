"""GPU, 2 ranks (NCCL): the data-parallel engine averages gradients correctly -- two ranks fed DIFFERENT halves of a batch
end up with the same parameters as one rank fed the whole batch (mean-reduced loss => gradient of the full batch is the
average of the half-batch gradients).  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import build_model

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _deterministic_attention_backward():
    """Engine-vs-engine equality is asserted to ~1e-5 after Adam steps: that needs the deterministic two-kernel attention
    backward (the default one-pass kernel reduces dQ with L2 atomics: reproducible to fp32 rounding, not bit for bit).
    ops.attn_bwd reads the variable at call time; spawned ranks inherit it."""
    old = os.environ.get("O2_ATTN_BWD_TWO_PASS")
    os.environ["O2_ATTN_BWD_TWO_PASS"] = "1"
    yield
    if old is None:
        os.environ.pop("O2_ATTN_BWD_TWO_PASS", None)
    else:
        os.environ["O2_ATTN_BWD_TWO_PASS"] = old


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _make(cfg, sd, dev, shard=False, dtype=torch.float32, full_shard=False, ckpt=False):
    from orbit2_b200 import engine, losses
    m = build_model(cfg, sd, dev, dtype)
    m.activation_checkpointing = ckpt
    loss = losses.METRICS_REGISTRY["mse"](aggregate_only=True, metainfo=losses.MetricsMetaInfo(cfg["in_vars"], cfg["out_vars"], None, None))
    return engine.TrainEngine(m, loss, cfg["in_vars"], cfg["out_vars"], cfg["var_weights"], lr=1e-3, betas=(0.9, 0.99), weight_decay=1e-5,
                              shard_optimizer=shard, shard_params=full_shard)


def _same_params(a, b, dtype):
    """Engine b (FULL_SHARD) holds the same parameters as engine a (replicated): fp32 masters of the sharded weights
    gathered from the shards, the replicated rest straight from the flat buffer."""
    sa, sb = a.full_state_dict(), b.full_state_dict()
    tol = 1e-5 if dtype == torch.float32 else 2e-3        # bf16 arm: split-K atomics order + bf16 operands
    bad = {k: (sa[k] - sb[k]).abs().max().item() for k in sa
           if (sa[k] - sb[k]).abs().max().item() > tol * sa[k].abs().max().item() + 1e-7}
    return bad


@pytest.mark.parametrize("dtype,ckpt", [(torch.float32, False), (torch.bfloat16, True)])
def test_full_shard_world1_equals_replicated(dtype, ckpt, tmp_path):
    """FULL_SHARD engine on one GPU (gathers are copies): the lookup-driven unit hand-over, the two rotating weight /
    gradient slots and the per-shard AdamW reproduce the replicated engine step for step."""
    from oracle import cases, reslim_oracle as O
    cfg = cases.get_case("tiny")
    sd = O.init_state_dict(cfg, seed=9)
    x, y = O.synthetic_batch(cfg, 2, cfg["in_vars"], cfg["out_vars"], seed=9)
    a = _make(cfg, sd, "cuda:0", dtype=dtype, ckpt=ckpt)
    b = _make(cfg, sd, "cuda:0", dtype=dtype, full_shard=True, ckpt=ckpt)
    assert b.fs is not None and len(b.fs.units) == cfg["depth"] + 1
    assert b.model.blocks[0].mlp.fc1.weight.numel() == 0           # the module keeps no full copy of a sharded weight
    for _ in range(3):
        va = a.step(x.cuda(), y.cuda())
        vb = b.step(x.cuda(), y.cuda())
        assert abs(va[-1].item() - vb[-1].item()) <= (1e-5 if dtype == torch.float32 else 5e-3) * abs(va[-1].item())
    bad = _same_params(a, b, dtype)
    assert not bad, bad
    assert _ckpt_roundtrip(b, lambda: _make(cfg, sd, "cuda:0", dtype=dtype, full_shard=True, ckpt=ckpt), x.cuda(), y.cuda(),
                           str(tmp_path))


def _ckpt_roundtrip(eng, make, x, y, ckdir):
    """save_checkpoint (collective gathers in the sharded modes, rank 0 writes) -> a fresh engine of the same mode resumes
    and takes the same next step: weights, Adam moments, step count and bias correction all survived."""
    from orbit2_b200 import trainer
    path = os.path.join(ckdir, f"ck_{eng.sharded}_{eng.fs is not None}_{eng.act}.ckpt")
    trainer.save_checkpoint(path, 4, eng, {"last_epoch": 5})
    if dist.is_initialized():
        dist.barrier()
    ck = torch.load(path, map_location="cpu", weights_only=False)
    ok = ck["optimizer_state_dict"] is not None and len(ck["optimizer_state_dict"]["state"]) > 0
    opt = torch.optim.AdamW([torch.nn.Parameter(v.clone()) for v in ck["model_state_dict"].values()], lr=1.0)
    opt.load_state_dict(ck["optimizer_state_dict"])        # the reference's optimizer accepts it (param order + shapes)
    eng2 = make()
    ok = ok and trainer.load_checkpoint(path, eng2) == 5 and eng2.step_count == eng.step_count
    va, vb = eng.step(x, y), eng2.step(x, y)
    torch.cuda.synchronize()
    sa, sb = eng.full_state_dict(), eng2.full_state_dict()
    tol = 1e-6 if eng.act == torch.float32 else 2e-3
    bad = [k for k in sa if (sa[k] - sb[k]).abs().max().item() > tol * sa[k].abs().max().item() + 1e-8]
    return ok and not bad and abs(va[-1].item() - vb[-1].item()) <= 5e-3 * abs(va[-1].item())


def _worker(rank, world, port, ret, ckdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import cases, reslim_oracle as O
    cfg = cases.get_case("tiny")
    sd = O.init_state_dict(cfg, seed=9)
    x, y = O.synthetic_batch(cfg, 4, cfg["in_vars"], cfg["out_vars"], seed=9)
    eng = _make(cfg, sd, f"cuda:{rank}")
    assert eng.world == 2
    for _ in range(3):
        eng.step(x[2 * rank:2 * rank + 2].cuda(), y[2 * rank:2 * rank + 2].cuda())
    torch.cuda.synchronize()
    ret[rank] = eng.flat_p.detach().cpu()
    # FSDP-style sharded optimizer (reduce-scatter + sharded AdamW + all-gather) must give the same parameters
    for dtype in (torch.float32, torch.bfloat16):
        a = _make(cfg, sd, f"cuda:{rank}", shard=False, dtype=dtype)
        b = _make(cfg, sd, f"cuda:{rank}", shard=True, dtype=dtype)
        assert b.sharded and b.flat_m.numel() == b.shard and b.shard * 2 >= a.total
        for _ in range(3):
            xa, ya = x[2 * rank:2 * rank + 2].cuda(), y[2 * rank:2 * rank + 2].cuda()
            a.step(xa, ya)
            b.step(xa, ya)
        torch.cuda.synchronize()
        n = a.total
        ret[f"shard{rank}{dtype}"] = bool(torch.allclose(a.flat_p[:n], b.flat_p[:n], rtol=1e-5, atol=1e-7)) and \
            (a.flat_b is None or bool(torch.equal(a.flat_b[:n], b.flat_b[:n]) or
                                      (a.flat_b[:n].float() - b.flat_b[:n].float()).abs().max().item() < 1e-2))
        ret[f"shard-ckpt{rank}{dtype}"] = _ckpt_roundtrip(b, lambda: _make(cfg, sd, f"cuda:{rank}", shard=True, dtype=dtype),
                                                          x[2 * rank:2 * rank + 2].cuda(), y[2 * rank:2 * rank + 2].cuda(), ckdir)
    # FULL_SHARD (per-Block all-gather / reduce-scatter, sharded master + Adam) against plain data parallel
    for dtype, ckpt in ((torch.float32, False), (torch.bfloat16, True)):
        a = _make(cfg, sd, f"cuda:{rank}", dtype=dtype, ckpt=ckpt)
        b = _make(cfg, sd, f"cuda:{rank}", dtype=dtype, full_shard=True, ckpt=ckpt)
        assert b.fs.world == 2 and b.fs.master[1].numel() * 2 == b.fs.S[1] * 2
        for _ in range(3):
            xa, ya = x[2 * rank:2 * rank + 2].cuda(), y[2 * rank:2 * rank + 2].cuda()
            a.step(xa, ya)
            b.step(xa, ya)
        torch.cuda.synchronize()
        bad = _same_params(a, b, dtype)
        ret[f"shard-full{rank}{dtype}"] = not bad
        ret[f"shard-full-ckpt{rank}{dtype}"] = _ckpt_roundtrip(
            b, lambda: _make(cfg, sd, f"cuda:{rank}", dtype=dtype, full_shard=True, ckpt=ckpt),
            x[2 * rank:2 * rank + 2].cuda(), y[2 * rank:2 * rank + 2].cuda(), ckdir)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_equal_one_rank_full_batch(tmp_path):
    from oracle import cases, reslim_oracle as O
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret, str(tmp_path)), nprocs=2, join=True)
    cfg = cases.get_case("tiny")
    sd = O.init_state_dict(cfg, seed=9)
    x, y = O.synthetic_batch(cfg, 4, cfg["in_vars"], cfg["out_vars"], seed=9)
    eng = _make(cfg, sd, "cuda:0")
    for _ in range(3):
        eng.step(x.cuda(), y.cuda())
    ref = eng.flat_p.detach().cpu()
    assert all(v for k, v in ret.items() if str(k).startswith("shard")), dict((k, v) for k, v in ret.items() if str(k).startswith("shard"))
    assert torch.equal(ret[0], ret[1])                                  # replicas stay bit-identical
    assert (ret[0] - ref).abs().max().item() < 2e-5 * ref.abs().max().item() + 1e-7
