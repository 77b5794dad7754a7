"""Multi-GPU TILES inference demo / timing: the low-resolution field is sharded over a div_v x div_h rank grid, halos are
exchanged over NCCL, every rank runs the network on its tile, the inner output blocks are all-gathered.
    torchrun --nproc-per-node 4 tools/tiles_infer.py --div-v 2 --div-h 2 --overlap 4 [--workload 117m]
Rank 0 checks the stitched result against the same tiles run sequentially on one GPU (tiles.tiled_forward)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import cases, reslim_oracle as O  # noqa: E402  (synthetic inputs only)
from orbit2_b200 import tiles  # noqa: E402
from orbit2_b200.reslim import Res_Slim_ViT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--div-v", type=int, default=2)
ap.add_argument("--div-h", type=int, default=2)
ap.add_argument("--overlap", type=int, default=4)
ap.add_argument("--workload", default="117m")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--field", type=int, nargs=2, default=None, metavar=("H", "W"),
                help="low-resolution field size instead of the workload's grid, e.g. 896 1440 -> a 3584 x 5760 output: the "
                     "size class of an 800 m CONUS grid (the reference's DAYMET grids are not in its repository)")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
assert world == a.div_v * a.div_h, "one rank per tile"
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = cases.get_case(a.workload)
if a.field is not None:
    cfg = dict(cfg, img_size=list(a.field))
H, W = cfg["img_size"]
th, tw = tiles.check_tiling(H, W, a.div_v, a.div_h, a.overlap, cfg["patch_size"])
torch.manual_seed(0)
m = Res_Slim_ViT(cfg["default_vars"], cases.get_case(a.workload)["img_size"], len(cfg["default_vars"]), cfg["out_channels"], 1,
                 patch_size=cfg["patch_size"], drop_path=0.0, drop_rate=0.0, learn_pos_emb=True, embed_dim=cfg["embed_dim"],
                 depth=cfg["depth"], decoder_depth=cfg["decoder_depth"], num_heads=cfg["num_heads"],
                 compute_dtype=torch.bfloat16)
with torch.no_grad():
    m.var_embed.normal_(0, 0.02); m.var_query.normal_(0, 0.02)
m.spatial_resolution = cfg["spatial_resolution"]
m.img_size = (th, tw)
m = m.cuda().eval()
x, _ = O.synthetic_batch(cfg, a.batch, cfg["in_vars"], cfg["out_vars"], seed=0)      # same global field on every rank
geo = tiles.ShardedField(H, W, a.div_v, a.div_h, a.overlap)
oy1, oy2, ox1, ox2 = geo.own(rank)
x_own = x[:, :, oy1:oy2, ox1:ox2].contiguous().cuda()
with torch.no_grad():
    for it in range(a.iters + 1):
        if it == 1:
            torch.cuda.synchronize(); dist.barrier()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        blk = tiles.sharded_tiled_forward(m, x_own, cfg["in_vars"], cfg["out_vars"], geo, rank)
        full = tiles.gather_output(blk, geo)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.iters], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        seq = tiles.tiled_forward(m, x.cuda(), cfg["in_vars"], cfg["out_vars"], a.div_v, a.overlap, div_h=a.div_h)
        torch.cuda.synchronize()
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True); s0.record()
        seq = tiles.tiled_forward(m, x.cuda(), cfg["in_vars"], cfg["out_vars"], a.div_v, a.overlap, div_h=a.div_h)
        s1.record(); torch.cuda.synchronize()
        err = (full.float() - seq.float()).abs().max().item() / seq.float().abs().max().item()
        print(f"TILES {a.div_v}x{a.div_h} overlap {a.overlap}: tile {th}x{tw}, halo {geo.halo_bytes(0, len(cfg['in_vars']), a.batch)} B/rank, "
              f"{ms.item():.2f} ms per field (max over ranks), {a.batch * 1e3 / ms.item():.2f} fields/s, "
              f"stitched vs sequential rel err {err:.2e}; the same tiles one after the other on one GPU: {s0.elapsed_time(s1):.2f} ms; "
              f"field {H}x{W} -> {H * cfg['superres_mag']}x{W * cfg['superres_mag']}, {th // cfg['patch_size'] * (tw // cfg['patch_size'])} tokens per tile")
        assert err < 1e-6
dist.destroy_process_group()
