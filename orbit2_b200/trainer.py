"""torchrun-compatible re-host of the reference training driver (examples/intermediate_downscaling.py) for the hot path.

Consumes the reference YAMLs unchanged (configs/interm_*.yaml: sections trainer / parallelism / tiling / model / data, keys
read at intermediate_downscaling.py:393-448), builds the drop-in model + registry loss + TrainEngine, runs the epoch loop
with the per-epoch warmup-cosine schedule (models/lr_scheduler.py:9-115, stepped once per epoch at :756) and writes
checkpoints in the reference's format (:775-791: epoch, model_state_dict, optimizer_state_dict, scheduler_state_dict;
the optimizer state is emitted in torch.optim.AdamW's own layout so either side can resume the other's run).

Batches come from any iterable yielding ``(x, y, in_variables, out_variables)`` like the reference's collate
(itermodule.py:451-469): ``--synthetic H W`` uses seeded random fields of that grid; ``--npz`` reads the reference's shard
directories named by the YAML (``data.low_res_dir`` / ``high_res_dir``) through ``orbit2_b200.data`` (raw fields copied to
the GPU, normalised there) and ends every epoch with the validation metrics of the downscaling task (rmse, pearson,
mean_bias on denormalised fields, utils/loaders.py:251-252).  Environment comes from torchrun (RANK / LOCAL_RANK /
WORLD_SIZE), not SLURM.

    torchrun --nproc-per-node 8 -m orbit2_b200.trainer /path/to/configs/interm_117m.yaml --data-key ERA5_2 \
             --synthetic 180 360 --steps-per-epoch 20 --epochs 2
"""
from __future__ import annotations

import argparse
import math
import os
import time
from typing import Dict, Iterable, Optional, Tuple

import torch
import torch.distributed as dist

from . import losses
from .engine import GradScaler, TrainEngine
from .reslim import Res_Slim_ViT


def load_config(path: str) -> dict:
    import yaml
    with open(path) as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def model_kwargs(conf: dict) -> dict:
    m = conf["model"]
    return dict(superres_mag=m["superres_mag"], cnn_ratio=m["cnn_ratio"], patch_size=m["patch_size"],
                embed_dim=m["embed_dim"], depth=m["depth"], decoder_depth=m["decoder_depth"], num_heads=m["num_heads"],
                mlp_ratio=m["mlp_ratio"], drop_path=m["drop_path"], drop_rate=m["drop_rate"])


def warmup_cosine_lr(epoch: int, base_lr: float, warmup_epochs: int, max_epochs: int, warmup_start_lr: float = 0.0,
                     eta_min: float = 0.0) -> float:
    """Closed form of LinearWarmupCosineAnnealingLR (lr_scheduler.py:97-115) = value after `epoch` scheduler steps."""
    if epoch < warmup_epochs:
        return warmup_start_lr + epoch * (base_lr - warmup_start_lr) / max(1, warmup_epochs - 1)
    return eta_min + 0.5 * (base_lr - eta_min) * (1 + math.cos(math.pi * (epoch - warmup_epochs) / (max_epochs - warmup_epochs)))


def build(conf: dict, data_key: str, img_size: Tuple[int, int], device, pos_grid: Optional[Tuple[int, int]] = None):
    """-> (model, loss, engine).  img_size = the low-resolution grid of this dataset (load_architecture reads it from
    data_module.get_data_dims(), loaders.py:261); spatial_resolution / variable lists come from the YAML's data section."""
    d, t = conf["data"], conf["trainer"]
    default_vars = list(d["default_vars"])
    in_vars = list(d["dict_in_variables"][data_key])
    out_vars = list(d["dict_out_variables"][data_key])
    dtype = torch.bfloat16 if t["data_type"] == "bfloat16" else torch.float32
    if conf["parallelism"]["tensor_par"] != 1 and int(os.environ.get("O2_IGNORE_TP", "1")):
        pass        # the reference's 1b/10b YAMLs ask for TP=4 on Frontier; one 8-GPU B200 box runs them data-parallel
    model = Res_Slim_ViT(default_vars, pos_grid or img_size, len(default_vars), len(out_vars), 1, learn_pos_emb=True,
                         compute_dtype=dtype, **model_kwargs(conf))
    model.spatial_resolution = d["spatial_resolution"][data_key]
    model.img_size = tuple(img_size)
    # the reference wraps every Block in a checkpoint wrapper (intermediate_downscaling.py:583-590, 635-637); with 180 GB
    # per GPU the 117M / 1B activations fit, so recomputation is opt-in here (O2_ACT_CKPT=1 or trainer --act-ckpt)
    model.activation_checkpointing = bool(int(os.environ.get("O2_ACT_CKPT", "0")))
    model = model.to(device)
    meta = losses.MetricsMetaInfo(in_vars, out_vars, None, None)
    loss = losses.METRICS_REGISTRY[t["train_loss"]](aggregate_only=True, metainfo=meta)
    m = conf["model"]
    full_shard = bool(int(os.environ.get("O2_FULL_SHARD", "0")))
    eng = TrainEngine(model, loss, in_vars, out_vars, d["var_weights"], lr=float(m["lr"]),
                      betas=(float(m["beta_1"]), float(m["beta_2"])), weight_decay=float(m["weight_decay"]),
                      # fsdp > 1 in the YAML -> sharded Adam state + update; O2_FULL_SHARD=1 / trainer --full-shard ->
                      # FSDP FULL_SHARD (per-Block weight shards, needed for interm_10b: intermediate_downscaling.py:615-617)
                      shard_optimizer=int(conf["parallelism"].get("fsdp", 1)) > 1 and not full_shard,
                      shard_params=full_shard,
                      # bf16 branch of the reference: ShardedGradScaler(init_scale=8192, growth_interval=100), floor 128
                      # (intermediate_downscaling.py:493-495, 733-742)
                      grad_scaler=GradScaler() if (dtype == torch.bfloat16 and not int(os.environ.get("O2_GRAPH", "0")))
                      else None)
    # launch-bound configurations (interm_8m): replay the step as one CUDA graph (trainer --graph / O2_GRAPH=1); only
    # where the engine allows it: one GPU, replicated parameters; dropout / drop-path stay on (device-resident step word),
    # the dynamic grad scaler is left out (its skip decision is taken on the host; bf16 has fp32's exponent range)
    if int(os.environ.get("O2_GRAPH", "0")):
        eng.enable_graph()
    return model, loss, eng


# ------------------------------------------------------------------------------------------------ checkpoints
def _adam_moments(eng: TrainEngine) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """name -> (exp_avg, exp_avg_sq) on the host for EVERY trained parameter, whatever the sharding mode.  Collective
    when the Adam state is sharded (``shard_optimizer``: one all-gather of the flat m / v; FULL_SHARD: one per unit)."""
    if eng.sharded:
        m = torch.empty(eng.shard * eng.world, device=eng.device, dtype=torch.float32)
        v = torch.empty_like(m)
        dist.all_gather_into_tensor(m, eng.flat_m, group=eng.pg)
        dist.all_gather_into_tensor(v, eng.flat_v, group=eng.pg)
    else:
        m, v = eng.flat_m, eng.flat_v
    out = {}
    for n in eng.names:
        if n in eng.frozen:
            continue
        lo, hi = eng.layout.range[n]
        shape = eng.P[n].shape
        out[n] = (m[lo:hi].view(shape).cpu().clone(), v[lo:hi].view(shape).cpu().clone())
    if eng.fs is not None:
        out.update(eng.fs.full_adam_state())
    return out


def optimizer_state_dict(eng: TrainEngine) -> dict:
    """torch.optim.AdamW.state_dict() layout (parameter ids in ``model.named_parameters()`` order, like an optimizer
    built from ``model.parameters()``: intermediate_downscaling.py:642-644) for the engine's Adam state.  Collective in
    the sharded modes: every rank calls it, the result is complete on every rank."""
    order = [n for n, _ in eng.model.named_parameters()]
    mom = _adam_moments(eng) if eng.step_count > 0 else {}
    state = {}
    for i, n in enumerate(order):
        if n in mom:
            state[i] = {"step": torch.tensor(float(eng.step_count)), "exp_avg": mom[n][0], "exp_avg_sq": mom[n][1]}
    group = {"lr": eng.lr, "betas": tuple(eng.betas), "eps": eng.eps, "weight_decay": eng.weight_decay, "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "params": list(range(len(order)))}
    return {"state": state, "param_groups": [group]}


def load_optimizer_state_dict(eng: TrainEngine, sd: dict):
    g = sd["param_groups"][0]
    eng.lr, eng.betas, eng.eps, eng.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
    steps = [int(float(s["step"])) for s in sd["state"].values()]
    eng.step_count = max(steps) if steps else 0
    order = [n for n, _ in eng.model.named_parameters()]
    o0, o1 = eng.own
    for i, st in sd["state"].items():
        n = order[int(i)]
        if eng.fs is not None and n in eng.fs.where:
            eng.fs.load_tensor(n, exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"])
            continue
        lo, hi = eng.layout.range[n]
        a, b = max(lo, o0), min(hi, o1)                    # sharded optimizer: this rank keeps its slice of the moments
        if a < b:
            eng.flat_m[a - o0:b - o0].copy_(st["exp_avg"].reshape(-1)[a - lo:b - lo])
            eng.flat_v[a - o0:b - o0].copy_(st["exp_avg_sq"].reshape(-1)[a - lo:b - lo])


def save_checkpoint(path: str, epoch: int, eng: TrainEngine, sched: dict):
    """The reference's checkpoint (intermediate_downscaling.py:775-791).  EVERY rank calls this (the sharded modes gather
    the weights and the Adam moments collectively); rank 0 writes the file."""
    ck = {"epoch": epoch, "model_state_dict": {k: v.detach().cpu() for k, v in eng.full_state_dict().items()},
          "optimizer_state_dict": optimizer_state_dict(eng), "scheduler_state_dict": dict(sched)}
    if not dist.is_initialized() or dist.get_rank() == 0:
        torch.save(ck, path)


def load_checkpoint(path: str, eng: TrainEngine) -> int:
    """Returns the epoch to start from (reference: epoch_start = ckpt['epoch'] + 1, intermediate_downscaling.py:659-672)."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    with torch.no_grad():
        sd = ck["model_state_dict"]
        for n, p in eng.model.named_parameters():
            if eng.fs is not None and n in eng.fs.where:
                eng.fs.load_tensor(n, value=sd[n])
            else:
                p.copy_(sd[n].to(p.device))                # parameters are views of the flat master buffer
        if eng.flat_b is not None:
            from . import ops
            ops.cast_bf16(eng.flat_p, eng.flat_b)
    if ck.get("optimizer_state_dict") is not None:         # model-only checkpoints (older sharded runs) resume the weights
        load_optimizer_state_dict(eng, ck["optimizer_state_dict"])
    if eng.fs is not None:
        eng.fs.after_step()                                # the gathered operand copies are stale
    return int(ck["epoch"]) + 1


def interpolate_pos_embed(pos_embed: torch.Tensor, patch_size: int, new_size: Tuple[int, int]) -> torch.Tensor:
    """Checkpoint pos_embed [1, L0, D] (stored grid h x 2h, the reference's fixed W / H = 2 assumption) -> the token grid
    of ``new_size`` (pixels) by bicubic resampling (components/pos_embed.py:73-99); unchanged when the row counts agree."""
    n, d = pos_embed.shape[-2], pos_embed.shape[-1]
    oh = int((n // 2) ** 0.5)
    ow = 2 * oh
    nh, nw = new_size[0] // patch_size, new_size[1] // patch_size
    if oh == nh:
        return pos_embed
    t = pos_embed.reshape(-1, oh, ow, d).permute(0, 3, 1, 2)
    t = torch.nn.functional.interpolate(t, size=(nh, nw), mode="bicubic", align_corners=False)
    return t.permute(0, 2, 3, 1).flatten(1, 2)


def load_pretrained_weights(model, state: Dict[str, torch.Tensor], log=print) -> Dict[str, list]:
    """``trainer.pretrain`` (examples/intermediate_downscaling.py:116-153): keys the model does not have are dropped, keys
    whose shape differs are dropped EXCEPT ``pos_embed``, which is resampled to the model's grid; the rest is copied in
    (``strict=False``).  Works on a plain module or on one whose parameters are views of an engine's flat buffer (copies
    are in place).  Returns {"loaded", "dropped", "missing"}."""
    own = model.state_dict()
    state = dict(state)
    dropped = []
    for k in list(state.keys()):
        if k not in own:
            log(f"Removing key {k} from pretrained checkpoint: no exist")
            dropped.append(k); del state[k]
        elif state[k].shape != own[k].shape:
            if k == "pos_embed":
                state[k] = interpolate_pos_embed(state[k], model.patch_size, tuple(model.img_size))
                if state[k].shape != own[k].shape:
                    dropped.append(k); del state[k]
            else:
                log(f"Removing key {k} from pretrained checkpoint: no matching shape {tuple(state[k].shape)} {tuple(own[k].shape)}")
                dropped.append(k); del state[k]
    with torch.no_grad():
        for k, v in state.items():
            own[k].copy_(v.to(own[k].device, own[k].dtype))
    return {"loaded": sorted(state), "dropped": dropped, "missing": sorted(set(own) - set(state))}


def load_pretrain_checkpoint(path: str, eng: TrainEngine, log=print):
    ck = torch.load(path, map_location="cpu", weights_only=False)
    rep = load_pretrained_weights(eng.model, ck["model_state_dict"], log)
    if eng.flat_b is not None:                             # refresh the bf16 operand copy of the flat master buffer
        from . import ops
        ops.cast_bf16(eng.flat_p, eng.flat_b)
    return rep


# ------------------------------------------------------------------------------------------------ loop
def synthetic_loader(conf, data_key, grid, batch, steps, device, seed):
    d = conf["data"]
    in_vars, out_vars = list(d["dict_in_variables"][data_key]), list(d["dict_out_variables"][data_key])
    mag = conf["model"]["superres_mag"]
    g = torch.Generator(device="cpu").manual_seed(seed)
    for _ in range(steps):
        x = torch.randn(batch, len(in_vars), grid[0], grid[1], generator=g)
        y = torch.randn(batch, len(out_vars), grid[0] * mag, grid[1] * mag, generator=g)
        if losses.PRECIP in in_vars:
            i = in_vars.index(losses.PRECIP)
            x[:, i] = torch.log1p(torch.relu(x[:, i]) * 2.0)
        if losses.PRECIP in out_vars:
            i = out_vars.index(losses.PRECIP)
            y[:, i] = torch.log1p(torch.relu(y[:, i]) * 2.0)
        yield x.pin_memory().to(device, non_blocking=True), y.pin_memory().to(device, non_blocking=True), in_vars, out_vars


def npz_data(conf: dict, data_key: str, device, rank: int, world: int):
    """DownscalingData over the YAML's shard directories (intermediate_downscaling.py:453-476 builds the same module)."""
    from .data import DownscalingData
    d, t = conf["data"], conf["trainer"]
    til = conf.get("tiling", {}) or {}
    div = int(til.get("div", 1)) if til.get("do_tiling", False) else 1
    return DownscalingData(d["low_res_dir"][data_key], d["high_res_dir"][data_key], list(d["dict_in_variables"][data_key]),
                           list(d["dict_out_variables"][data_key]), t["batch_size"], device, rank, world, div=div,
                           overlap=int(til.get("overlap", 4)), subsample=1, buffer_size=int(t.get("buffer_size", 0)), seed=rank)


def validate(eng: TrainEngine, dm, max_batches: Optional[int] = None) -> Dict[str, float]:
    """One pass over the validation split: clip_replace_constant, then rmse / pearson / mean_bias on denormalised fields
    (intermediate_downscaling.py:321-371), averaged over batches and ranks."""
    from .data import denorm_affine
    model = eng.model
    was_training = model.training
    model.eval()
    lat, lon = dm.get_lat_lon()
    meta = losses.MetricsMetaInfo(dm.in_vars, dm.out_vars, lat, lon)
    mets = {n: losses.METRICS_REGISTRY[n](aggregate_only=True, metainfo=meta, denorm=denorm_affine(dm.out_stats))
            for n in ("rmse", "pearson", "mean_bias")}
    tot = {n: torch.zeros((), device="cuda") for n in mets}
    k = 0
    with torch.no_grad():
        for x, y, in_vars, out_vars in dm.loader("val"):
            pred = model(x, in_vars, out_vars)
            yc = y[:, :, :pred.shape[2], :pred.shape[3]].contiguous()
            pred = losses.clip_replace_constant(yc, pred, out_vars)
            for n, fn in mets.items():
                tot[n] += fn(pred, yc)
            k += 1
            if max_batches and k >= max_batches:
                break
    out = {}
    for n in mets:
        v = tot[n] / max(k, 1)
        if dist.is_initialized():
            dist.all_reduce(v, op=dist.ReduceOp.AVG)
        out[n] = float(v)
    model.train(was_training)
    return out


def train(conf: dict, data_key: str, grid, epochs: int, steps_per_epoch: int, ckpt_dir: Optional[str] = None,
          resume: Optional[str] = None, loader_factory=None, log=print, data_module=None):
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    device = torch.device("cuda", torch.cuda.current_device())
    model, loss, eng = build(conf, data_key, grid, device)
    m = conf["model"]
    sched = dict(warmup_epochs=m["warmup_epochs"], max_epochs=conf["trainer"]["max_epochs"],
                 warmup_start_lr=float(m["warmup_start_lr"]), eta_min=float(m["eta_min"]), base_lr=float(m["lr"]))
    pre = conf["trainer"].get("pretrain")
    if pre and not resume:                                 # like the reference: a resumed checkpoint wins over the pretrain path
        if not os.path.exists(pre):
            raise FileNotFoundError(f"trainer.pretrain = {pre}: pretrain path does not exist")
        rep = load_pretrain_checkpoint(pre, eng, log if rank == 0 else (lambda *_: None))
        if rank == 0:
            log(f"pretrained weights: {len(rep['loaded'])} loaded, {len(rep['dropped'])} dropped, {len(rep['missing'])} kept from init")
    epoch0 = load_checkpoint(resume, eng) if resume else 0
    B = conf["trainer"]["batch_size"]
    hist = []
    for epoch in range(epoch0, epoch0 + epochs):
        eng.lr = warmup_cosine_lr(epoch, sched["base_lr"], sched["warmup_epochs"], sched["max_epochs"],
                                  sched["warmup_start_lr"], sched["eta_min"])
        if data_module is not None:
            loader = data_module.loader("train")
        else:
            loader = (loader_factory or synthetic_loader)(conf, data_key, grid, B, steps_per_epoch, device, 1000 * epoch + rank)
        t0 = time.perf_counter()
        tot, n = torch.zeros((), device=device), 0
        for x, y, in_vars, out_vars in loader:
            vec = eng.step(x, y)
            tot += vec[-1]
            n += 1
            if data_module is not None and steps_per_epoch and n >= steps_per_epoch:
                break
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        mean = (tot / max(n, 1)).item()
        hist.append(mean)
        val = validate(eng, data_module) if data_module is not None else None
        if rank == 0:
            log(f"epoch {epoch} lr {eng.lr:.3e} loss {mean:.5f} {n * B * world / dt:.2f} samples/s"
                + ("" if val is None else " val " + " ".join(f"{k} {v:.5f}" for k, v in val.items())), flush=True)
        if ckpt_dir:                                       # every rank: the sharded modes gather collectively
            os.makedirs(ckpt_dir, exist_ok=True)
            save_checkpoint(os.path.join(ckpt_dir, f"interm_epoch_{epoch}.ckpt"), epoch, eng, dict(sched, last_epoch=epoch + 1))
        if dist.is_initialized():
            dist.barrier()
    return hist, eng


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--data-key", default="ERA5_2")
    ap.add_argument("--synthetic", type=int, nargs=2, metavar=("H", "W"), default=None)
    ap.add_argument("--npz", action="store_true", help="read data.low_res_dir / high_res_dir[data-key] (reference shard layout)")
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--steps-per-epoch", type=int, default=10)
    ap.add_argument("--checkpoint-dir", default=None)
    ap.add_argument("--resume", default=None)
    ap.add_argument("--act-ckpt", action="store_true", help="per-Block activation recomputation (the reference's default)")
    ap.add_argument("--full-shard", action="store_true", help="FSDP FULL_SHARD: per-Block weight / gradient / Adam shards "
                    "(interm_1b / interm_10b); default is the YAML's parallelism.fsdp -> sharded optimizer")
    ap.add_argument("--graph", action="store_true", help="replay the training step as one captured CUDA graph (one GPU, "
                    "no dynamic grad scaler; for launch-bound configurations such as interm_8m, at the YAML's dropout)")
    a = ap.parse_args()
    if a.act_ckpt:
        os.environ["O2_ACT_CKPT"] = "1"
    if a.graph:
        os.environ["O2_GRAPH"] = "1"
    if a.full_shard:
        os.environ["O2_FULL_SHARD"] = "1"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    conf = load_config(a.config)
    if a.npz:
        dm = npz_data(conf, a.data_key, torch.device("cuda", local), int(os.environ.get("RANK", "0")),
                      int(os.environ.get("WORLD_SIZE", "1")))
        grid = dm.get_data_dims()[0][2:]
        train(conf, a.data_key, tuple(grid), a.epochs, a.steps_per_epoch, a.checkpoint_dir, a.resume, data_module=dm)
    else:
        if a.synthetic is None:
            ap.error("give --synthetic H W or --npz")
        train(conf, a.data_key, tuple(a.synthetic), a.epochs, a.steps_per_epoch, a.checkpoint_dir, a.resume)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
