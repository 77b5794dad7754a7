#!/usr/bin/env python
"""bench.py -- train samples/s of the ORBIT-2 Reslim hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 117m|8m] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One step = forward -> clip_replace_constant -> bayesian_tv loss (the shipped ``train_loss``, configs/interm_117m.yaml:10)
-> backward -> data-parallel gradient all-reduce -> AdamW, on one batch of synthetic ERA5-shaped fields
(interm_117m on the 180x360 -> 720x1440 grid, B = 8 per GPU: configs/interm_117m.yaml:6; weak scaling).
Prints ONE JSON line (rank 0).  ``value`` is timed with the batch resident in HBM, ``e2e`` re-times the same steps
through the public API (Res_Slim_ViT.forward + METRICS_REGISTRY loss + backward + AdamW) with pinned host inputs copied
in and the loss read back every step.  ``--impl reference`` times the CPU restatement of the reference (oracle/, SDPA
attention = the reference's FusedAttn.DEFAULT path) on the host cores -- the Python reference itself cannot travel to
the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train samples/sec at 1/2/4/8 B200 (117M, ERA5 1.0°→0.25°); % bf16/HBM roofline"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe): one background
    `nvidia-smi -lms 200` process, parsed when the timed region ends."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
        except Exception:
            self.proc = None

    def summary(self):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[c.strip() for c in ln.split(",")] for ln in out.splitlines() if ln.count(",") >= 7]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def flops_per_sample(cfg):
    """counted dense FLOPs, forward+backward (GEMMs x3, attention x3.5)."""
    p = cfg["patch_size"]
    L = (cfg["img_size"][0] // p) * (cfg["img_size"][1] // p)
    D, depth, dec, C, mag = cfg["embed_dim"], cfg["depth"], cfg["decoder_depth"], cfg["out_channels"], cfg["superres_mag"]
    lin = depth * 24 * L * D * D + dec * 2 * L * D * D + 2 * L * D * C * (mag * p) ** 2 + 2 * L * D * D
    att = depth * 4 * L * L * D
    return 3.0 * lin + 3.5 * att, att / depth, L


def kernel_source_hash():
    """sha256 (16 hex digits) of the attention kernel sources: a committed ncu DRAM-traffic figure is only reported while
    the kernels it was captured from are unchanged."""
    import hashlib
    h = hashlib.sha256()
    for f in ("attn_tc.cu", "attn_bwd_fused.cuh", "common.cuh"):
        try:
            h.update(open(os.path.join(ROOT, "orbit2_b200", "csrc", f), "rb").read())
        except OSError:
            return None
    return h.hexdigest()[:16]


def committed_traffic(kernel, B, L, heads):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, tools/ncu_traffic.py)
    from profiles/traffic.json -- None unless the capture is of this shape AND of the current kernel sources."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = tr["kernels"].get(kernel)
        if e and tr.get("src_sha") == kernel_source_hash() and (e["B"], e["N"], e["heads"]) == (B, L, heads):
            return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_step_time(case_name, B, steps, warmup, threads, budget_s=None):
    """Seconds per step of the CPU restatement (fp32, SDPA attention) of forward+clip+loss+backward.  With ``budget_s`` the
    number of warm-up / timed steps is cut (never the grid) so that the whole call ends inside the budget; returns
    (seconds per step, cfg, timed steps, warm-up steps actually run)."""
    import torch
    from oracle import cases, reslim_oracle as O
    torch.set_num_threads(threads)
    O.USE_SDPA = True
    cfg = cases.get_case(case_name)
    sd = {k: v.requires_grad_(True) for k, v in O.init_state_dict(cfg, 0).items()}
    x, y = O.synthetic_batch(cfg, B, cfg["in_vars"], cfg["out_vars"], 0)

    def one():
        t0 = time.perf_counter()
        loss = O.training_step(sd, cfg, x, y, cfg["in_vars"], cfg["out_vars"], "bayesian_tv", cfg["var_weights"])
        loss.backward()
        for v in sd.values():
            v.grad = None
        return time.perf_counter() - t0

    t_begin = time.perf_counter()
    times, done_w = [], 0
    for _ in range(warmup):
        last = one()
        done_w += 1
        if budget_s is not None and (time.perf_counter() - t_begin) + last * (1 + steps) > budget_s:
            break                                   # the remaining warm-up steps would eat the timed ones
    for _ in range(steps):
        times.append(one())
        if budget_s is not None and (time.perf_counter() - t_begin) + times[-1] > budget_s:
            break
    return sum(times) / len(times), cfg, len(times), done_w


REF_CASE = {"117m": ("117m", 1), "117m_90x180": ("117m_90x180", 1), "8m": ("8m", 8), "1b": ("1b", 1), "10b": ("10b", 1),
            "10b_d2": ("10b_d2", 1)}


def sample_note(cfg, B):
    return f"B={B} on the full {cfg['img_size'][0]}x{cfg['img_size'][1]} grid of the workload (the GPU arm's grid)"


def run_reference(args):
    """CPU arm.  ALWAYS the GPU arm's grid (interm_117m: 180x360 -> 720x1440, L = 16200), B = 1 per step (a B = 8 step is
    8 x 26 s on 16 cores and samples/s does not depend on B on the host); when K + W steps do not fit ``--ref-budget``
    seconds the NUMBER of steps is cut and reported, the grid never changes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    case, B = REF_CASE[args.workload]
    sec, cfg, n_timed, n_warm = cpu_reference_step_time(case, B, args.steps, args.warmup, threads, budget_s=args.ref_budget)
    val = B / sec
    L = (cfg["img_size"][0] // cfg["patch_size"]) * (cfg["img_size"][1] // cfg["patch_size"])
    sample = ("CPU restatement of the reference (oracle/, fp32, SDPA attention = the reference's FusedAttn.DEFAULT path), "
              "forward+clip+bayesian_tv+backward, " + sample_note(cfg, B) +
              f", {n_timed} timed steps after {n_warm} warm-up (requested {args.steps}+{args.warmup}, budget {args.ref_budget:.0f} s)")
    H_out = cfg["img_size"][0] * cfg["superres_mag"]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": n_timed, "warmup": n_warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"interm_{args.workload} Res_Slim_ViT ERA5 {cfg['img_size'][0]}x{cfg['img_size'][1]} -> "
                                   f"{H_out}x{cfg['img_size'][1] * cfg['superres_mag']}, V={len(cfg['in_vars'])} in / "
                                   f"{len(cfg['out_vars'])} out vars, fwd+clip+bayesian_tv+bwd (host cores)",
                       "per_gpu_batch": B, "tokens_per_sample": L, "same_grid_as_gpu_arm": True},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from oracle import cases, reslim_oracle as O            # synthetic inputs / weights only (bench baseline leg)
    from orbit2_b200 import _lib, engine, losses, ops
    from orbit2_b200.reslim import Res_Slim_ViT

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load(require_device=True)

    cfg = cases.get_case(args.workload)
    B = args.batch or {"117m": 8, "1b": 8}.get(args.workload, 32)       # 8m / 10b: 32 per GPU (configs/interm_*.yaml:6)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    with torch.device(dev if args.workload.startswith("10b") else "cpu"):     # 9.5 B parameters are initialised on the GPU
        model = Res_Slim_ViT(cfg["default_vars"], cfg["img_size"], len(cfg["default_vars"]), cfg["out_channels"], 1,
                             superres_mag=cfg["superres_mag"], cnn_ratio=cfg["cnn_ratio"], patch_size=cfg["patch_size"],
                             drop_path=args.drop, drop_rate=args.drop, learn_pos_emb=True, embed_dim=cfg["embed_dim"],
                             depth=cfg["depth"], decoder_depth=cfg["decoder_depth"], num_heads=cfg["num_heads"],
                             mlp_ratio=cfg["mlp_ratio"], compute_dtype=dtype)
    with torch.no_grad():                                   # zeros would hide the front end
        model.var_embed.normal_(0, 0.02)
        model.var_query.normal_(0, 0.02)
    model.spatial_resolution = cfg["spatial_resolution"]
    model.activation_checkpointing = args.ckpt
    model = model.to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    H_out = cfg["img_size"][0] * cfg["superres_mag"]
    meta = losses.MetricsMetaInfo(cfg["in_vars"], cfg["out_vars"], None, None)
    loss_fn = losses.METRICS_REGISTRY["bayesian_tv"](aggregate_only=True, metainfo=meta)
    eng = engine.TrainEngine(model, loss_fn, cfg["in_vars"], cfg["out_vars"], cfg["var_weights"], lr=2e-4,
                             betas=(0.9, 0.99), weight_decay=1e-5, shard_optimizer=args.shard,
                             shard_params=args.full_shard)

    x_h, y_h = O.synthetic_batch(cfg, B, cfg["in_vars"], cfg["out_vars"], seed=rank)
    x_h, y_h = x_h.pin_memory(), y_h.pin_memory()
    x_d, y_d = x_h.to(dev), y_h.to(dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps

    # ---- device-resident arm, with per-kernel CUDA events (same stream) for the roofline
    def dev_step():
        return eng.step(x_d, y_d)

    for _ in range(args.warmup):
        dev_step()
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    ops.TIMERS = {}
    launches0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        vec = dev_step()
    e1.record()
    sync_all()
    clocks = sampler.summary()
    ms_step = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms_step], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = t.item()
    timers, ops.TIMERS = ops.TIMERS, None
    launches = (ops.LAUNCHES - launches0) // args.steps
    loss_val = float(vec[-1].item())
    eager_ms = None
    if args.graph:
        # launch-bound workloads (interm_8m): the same step captured once as a CUDA graph and replayed.  The per-kernel
        # events above come from the eager pass (events cannot be read back from a capture); `value` is the replayed step.
        eager_ms = ms_step
        eng.enable_graph(warm_steps=1)
        for _ in range(max(args.warmup, 1)):
            dev_step()
        sync_all()
        sampler = ClockSampler(local)
        sampler.start()
        e0.record()
        for _ in range(args.steps):
            vec = dev_step()
        e1.record()
        sync_all()
        clocks = sampler.summary()
        ms_step = e0.elapsed_time(e1) / args.steps
        loss_val = float(vec[-1].item())

    kern = {}
    for name, evs in timers.items():
        if name == "gemm_flops":
            continue
        kern[name] = sum(a.elapsed_time(b) for a, b in evs) / args.steps          # ms per step
    gemm_flops = sum(timers.get("gemm_flops", [])) / args.steps

    # ---- end-to-end arm: public API, pinned host inputs in, loss out, every step
    opt_state = {"loss": None}
    loss_pub = losses.METRICS_REGISTRY["bayesian_tv"](aggregate_only=True, metainfo=meta)

    def e2e_step():
        x = x_h.to(dev, non_blocking=True)
        y = y_h.to(dev, non_blocking=True)
        eng.flat_g.zero_()
        pred = model(x, cfg["in_vars"], cfg["out_vars"])
        loss = loss_pub(pred, y, var_names=cfg["out_vars"], var_weights=cfg["var_weights"],
                        clip_out_variables=cfg["out_vars"])
        loss.backward()
        if world > 1:
            eng.reducer.ready(eng.names)
            eng.reducer.finish()
        eng.optimizer_step()
        opt_state["loss"] = loss.item()                    # device -> host read of the step's result

    def e2e_step_engine():                                 # FULL_SHARD: the module holds no full weights, the engine step
        x = x_h.to(dev, non_blocking=True)                 # is the public call (trainer.py drives exactly this)
        y = y_h.to(dev, non_blocking=True)
        opt_state["loss"] = eng.step(x, y)[-1].item()

    if args.full_shard:
        e2e_step = e2e_step_engine
    model.train()
    ms_e2e = timed(e2e_step, args.steps, max(1, args.warmup // 2))

    peaks = load_peaks()
    fl_sample, attn_fwd_flops_blk, L = flops_per_sample(cfg)
    value = world * B / (ms_step * 1e-3)
    e2e_val = world * B / (ms_e2e * 1e-3)
    step_tflops = fl_sample * B / (ms_step * 1e-3) / 1e12

    # dominant kernel = the attention launch group with the largest share of the step
    ATT_MULT = {"attn_fwd": 1.0, "attn_bwd_dkv": 2.0, "attn_bwd_dq": 1.5, "attn_bwd_fused": 2.5}   # 2 / 4 / 3 / 5 GEMMs of 2*L*L*D
    att_names = [n for n in ATT_MULT if n in kern]
    roof = None
    dom = max(att_names, key=lambda n: kern[n]) if att_names else None
    if "gemm" in kern and (dom is None or kern["gemm"] > kern[dom]) and gemm_flops > 0:
        # GEMM-dominated workloads (interm_1b / 10b): the launch group is every o2_gemm call of the step
        nlaunch = (len(timers["gemm"]) / args.steps) or 1
        ach = gemm_flops / (kern["gemm"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_tc (all launches of the step)", "achieved": ach, "peak": peaks["tf_sust"],
                "unit": "TFLOP/s", "frac": ach / peaks["tf_sust"], "traffic": None,
                "peak_source": peaks["src"] + " (sustained cuBLAS bf16)", "avg_launch_ms": kern["gemm"] / nlaunch,
                "flops_per_launch": gemm_flops / nlaunch, "share_of_step": kern["gemm"] / ms_step}
    elif dom is not None:
        nlaunch = len(timers[dom]) / args.steps
        fl = attn_fwd_flops_blk * ATT_MULT[dom] * B                              # ALGORITHMIC FLOPs per launch
        avg_ms = kern[dom] / nlaunch
        ach = fl / (avg_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                "frac": ach / peaks["tf_sust"], "traffic": committed_traffic(dom, B, L, cfg["num_heads"]),
                "peak_source": peaks["src"] + " (sustained cuBLAS bf16)",
                "avg_launch_ms": avg_ms, "flops_per_launch": fl,
                "share_of_step": kern[dom] / ms_step}
    kernels_ms = {k: round(v, 3) for k, v in sorted(kern.items(), key=lambda kv: -kv[1])}
    if "gemm" in kern and kern["gemm"] > 0:
        kernels_ms["gemm_tflops"] = round(gemm_flops / (kern["gemm"] * 1e-3) / 1e12, 1)
    for n, mult in ATT_MULT.items():
        if n in kern:
            kernels_ms[n + "_tflops"] = round(attn_fwd_flops_blk * mult * B * cfg["depth"] / (kern[n] * 1e-3) / 1e12, 1)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and not args.workload.startswith("10b"):      # 9.5 B fp32 parameters + gradients: no host fit
        threads = os.cpu_count() or 1
        case, cb = REF_CASE[args.workload]
        sec, ccfg, nt, nw = cpu_reference_step_time(case, cb, 1, 1, threads)
        cpu = {"value": cb / sec, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"CPU restatement of the reference (oracle/, fp32, SDPA attention), {nt} fwd+clip+bayesian_tv+bwd "
                         f"step after {nw} warm-up, " + sample_note(ccfg, cb)}

    extra = None
    if world > 1 and not args.no_extra and args.workload == "117m" and not (args.full_shard or args.shard):
        # secondary, UNTIMED-by-the-headline measurements of the other two ways the path shards
        del eng, model, loss_fn, loss_pub, x_d, y_d
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        extra = extra_measurements(args, cfg, dev, rank, world, B, x_h, y_h, value)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"interm_{args.workload} Res_Slim_ViT ({n_params / 1e6:.1f}M params) ERA5 "
                                   f"{cfg['img_size'][0]}x{cfg['img_size'][1]} -> {H_out}x{cfg['img_size'][1] * cfg['superres_mag']}"
                                   f", V={len(cfg['in_vars'])} in / {len(cfg['out_vars'])} out vars, fwd+clip+bayesian_tv+bwd+allreduce+AdamW",
                       "per_gpu_batch": B, "global_batch": B * world, "tokens_per_sample": L, "parallelism": (f"fsdp{world} FULL_SHARD (per-Block all-gather fwd+bwd, gradient reduce-scatter, sharded fp32 master + Adam)" if args.full_shard else f"fsdp{world} (sharded Adam state + update, reduce-scatter / all-gather)" if (args.shard and world > 1) else f"dp{world}"),
                       "l2_policy": "inputs larger than L2 (activations of one step >> 126 MB), no explicit flush",
                       "dropout": args.drop, "activation_checkpointing": bool(args.ckpt),
                       **({"cuda_graph": True, "eager_ms_per_step": eager_ms} if args.graph else {})},
            "e2e": {"value": e2e_val, "unit": "samples/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": (x_h.numel() + y_h.numel()) * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "step_tflops": step_tflops, "step_frac_of_bf16_peak": step_tflops / peaks["tf_sust"],
            "kernels_ms_per_step": kernels_ms, "loss": loss_val,
            "hbm_peak_gib": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2),
        }
        if extra is not None:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extra_measurements(args, cfg, dev, rank, world, B, x_h, y_h, dp_value):
    """N > 1 only, after the headline region: (1) the same interm_117m step under FSDP FULL_SHARD (the mode interm_1b /
    interm_10b train in: per-Block all-gather in forward and backward, gradient reduce-scatter, sharded master + Adam;
    reference intermediate_downscaling.py:583-637) and (2) TILES inference of one synthetic CONUS-like field that is SHARDED
    over the ranks (div_v x div_h = world tiles, NCCL halo exchange of the raw input margins, per-tile network, all-gather of
    the inner output blocks; reference utils/visualize.py:125-311 runs the tiles one after the other on one rank), checked
    on rank 0 against the sequential tile loop.  Both are timed with CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    from oracle import reslim_oracle as O
    from orbit2_b200 import engine, losses, tiles
    from orbit2_b200.reslim import Res_Slim_ViT

    def make_model(tile_size=None):
        torch.manual_seed(0)
        m = Res_Slim_ViT(cfg["default_vars"], cfg["img_size"], len(cfg["default_vars"]), cfg["out_channels"], 1,
                         superres_mag=cfg["superres_mag"], cnn_ratio=cfg["cnn_ratio"], patch_size=cfg["patch_size"],
                         drop_path=0.0, drop_rate=0.0, learn_pos_emb=True, embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                         decoder_depth=cfg["decoder_depth"], num_heads=cfg["num_heads"], mlp_ratio=cfg["mlp_ratio"],
                         compute_dtype=torch.bfloat16)
        with torch.no_grad():
            m.var_embed.normal_(0, 0.02)
            m.var_query.normal_(0, 0.02)
        m.spatial_resolution = cfg["spatial_resolution"]
        if tile_size is not None:                # like the reference's data_config: pos_embed keeps its 2:1 grid and is
            m.img_size = tuple(tile_size)        # resampled (bicubic, pos_embed.py:103-138) to the tile grid on the fly
        return m.to(dev)

    def max_ms(ms):
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    out = {}
    # ---- (1) FULL_SHARD
    model = make_model()
    meta = losses.MetricsMetaInfo(cfg["in_vars"], cfg["out_vars"], None, None)
    eng = engine.TrainEngine(model, losses.METRICS_REGISTRY["bayesian_tv"](aggregate_only=True, metainfo=meta), cfg["in_vars"],
                             cfg["out_vars"], cfg["var_weights"], lr=2e-4, betas=(0.9, 0.99), weight_decay=1e-5,
                             shard_params=True)
    x_d, y_d = x_h.to(dev), y_h.to(dev)
    for _ in range(2):
        eng.step(x_d, y_d)
    dist.barrier()
    torch.cuda.synchronize()
    g0, s0 = eng.fs.gathered_elems, eng.fs.scattered_elems
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_fs = 3
    e0.record()
    for _ in range(n_fs):
        vec = eng.step(x_d, y_d)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = max_ms(e0.elapsed_time(e1) / n_fs)
    fs_val = world * B / (ms * 1e-3)
    out["full_shard_117m"] = {
        "value": fs_val, "unit": "samples/s", "ms_per_step": ms, "steps": n_fs, "vs_dp": fs_val / dp_value,
        "parallelism": f"fsdp{world} FULL_SHARD, one unit per Block",
        "gathered_bytes_per_step_per_rank": (eng.fs.gathered_elems - g0) // n_fs * 2,
        "reduce_scattered_bytes_per_step_per_rank": (eng.fs.scattered_elems - s0) // n_fs * 4,
        "loss": float(vec[-1].item())}
    del eng, model, x_d, y_d
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    # ---- (2) TILES over a sharded field: 360 x 720 low-resolution field -> 1440 x 2880 output, one tile per rank
    div_v, div_h = {2: (1, 2), 4: (2, 2), 8: (2, 4)}.get(world, (1, world))
    Hf, Wf, overlap = 360, 720, 4
    th, tw = tiles.check_tiling(Hf, Wf, div_v, div_h, overlap, cfg["patch_size"])
    model = make_model([th, tw]).eval()
    fcfg = dict(cfg, img_size=[Hf, Wf])
    xf, _ = O.synthetic_batch(fcfg, 1, cfg["in_vars"], cfg["out_vars"], seed=0)       # the same global field on every rank
    geo = tiles.ShardedField(Hf, Wf, div_v, div_h, overlap)
    oy1, oy2, ox1, ox2 = geo.own(rank)
    x_own = xf[:, :, oy1:oy2, ox1:ox2].contiguous().to(dev)
    n_it = 3
    with torch.no_grad():
        for it in range(n_it + 1):
            if it == 1:
                dist.barrier()
                torch.cuda.synchronize()
                e0.record()
            blk = tiles.sharded_tiled_forward(model, x_own, cfg["in_vars"], cfg["out_vars"], geo, rank)
            full = tiles.gather_output(blk, geo)
        e1.record()
        torch.cuda.synchronize()
        ms = max_ms(e0.elapsed_time(e1) / n_it)
        err, seq_ms = None, None
        if rank == 0:
            xg = xf.to(dev)
            seq = tiles.tiled_forward(model, xg, cfg["in_vars"], cfg["out_vars"], div_v, overlap, div_h=div_h)
            torch.cuda.synchronize()
            e0.record()
            seq = tiles.tiled_forward(model, xg, cfg["in_vars"], cfg["out_vars"], div_v, overlap, div_h=div_h)
            e1.record()
            torch.cuda.synchronize()
            seq_ms = e0.elapsed_time(e1)
            err = (full.float() - seq.float()).abs().max().item() / seq.float().abs().max().item()
    out["tiles_sharded_field"] = {
        "field": f"{Hf}x{Wf} -> {Hf * cfg['superres_mag']}x{Wf * cfg['superres_mag']}, V={len(cfg['in_vars'])}",
        "tiles": f"{div_v}x{div_h}", "tile": f"{th}x{tw}", "overlap": overlap,
        "halo_bytes_rank0": geo.halo_bytes(0, len(cfg["in_vars"]), 1), "ms_per_field": ms, "fields_per_s": 1e3 / ms,
        "sequential_one_gpu_ms": seq_ms, "stitched_vs_sequential_rel_err": err}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="117m", choices=["117m", "8m", "117m_90x180", "1b", "10b", "10b_d2"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--drop", type=float, default=0.0, help="drop_rate = drop_path (the reference YAMLs train at 0.1; the "
                    "headline and every parity run use 0)")
    ap.add_argument("--ckpt", action="store_true", help="per-Block activation recomputation (reference: checkpoint wrappers "
                    "on every Block under FSDP); the recomputed forward FLOPs are NOT counted in the roofline")
    ap.add_argument("--graph", action="store_true", help="replay the device-resident step as one captured CUDA graph "
                    "(single GPU, dropout 0; for launch-bound workloads such as 8m)")
    ap.add_argument("--shard", action="store_true", help="FSDP-style sharded optimizer instead of plain data parallel")
    ap.add_argument("--full-shard", action="store_true",
                    help="FSDP FULL_SHARD: GEMM weights, fp32 masters, gradients and Adam state sharded per Block "
                         "(all-gather in forward and backward, reduce-scatter of gradients)")
    ap.add_argument("--no-extra", action="store_true", help="N > 1: skip the secondary FULL_SHARD / sharded-field TILES "
                    "measurements appended under `extra`")
    ap.add_argument("--ref-budget", type=float, default=1500.0, help="--impl reference: seconds the CPU arm may take; the "
                    "number of steps is cut to fit, the grid is always the GPU arm's")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
