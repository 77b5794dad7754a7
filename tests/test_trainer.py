"""Trainer re-host: YAML schema, LR schedule (CPU); short run + checkpoint round trip (GPU)."""
import importlib.util
import math
import os

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF_SCHED = "/root/reference/src/climate_learn/models/lr_scheduler.py"


def test_own_yaml_has_reference_schema():
    from orbit2_b200 import trainer
    for name, dim in (("interm_8m", 256), ("interm_117m", 1024)):
        c = trainer.load_config(os.path.join(ROOT, "configs", name + ".yaml"))
        for sec, keys in (("trainer", ["max_epochs", "checkpoint", "pretrain", "batch_size", "num_workers", "buffer_size",
                                       "data_type", "train_loss"]),
                          ("parallelism", ["fsdp", "simple_ddp", "tensor_par", "seq_par"]),
                          ("tiling", ["do_tiling", "div", "overlap"]),
                          ("model", ["preset", "lr", "beta_1", "beta_2", "weight_decay", "warmup_epochs", "warmup_start_lr",
                                     "eta_min", "superres_mag", "cnn_ratio", "patch_size", "embed_dim", "depth",
                                     "decoder_depth", "num_heads", "mlp_ratio", "drop_path", "drop_rate"]),
                          ("data", ["low_res_dir", "high_res_dir", "spatial_resolution", "default_vars", "dict_in_variables",
                                    "dict_out_variables", "var_weights"])):
            for k in keys:
                assert k in c[sec], (name, sec, k)
        assert trainer.model_kwargs(c)["embed_dim"] == dim


@pytest.mark.reference
@pytest.mark.parametrize("cfg", ["interm_8m", "interm_117m", "interm_1b", "interm_10b"])
def test_reference_yaml_is_consumed_unchanged(cfg):
    from orbit2_b200 import trainer
    c = trainer.load_config(f"/root/reference/configs/{cfg}.yaml")
    kw = trainer.model_kwargs(c)
    assert kw["patch_size"] == 2 and kw["superres_mag"] == 4
    assert set(c["data"]["dict_out_variables"]["ERA5_2"]) <= set(c["data"]["dict_in_variables"]["ERA5_2"])


def test_lr_schedule_closed_form():
    from orbit2_b200.trainer import warmup_cosine_lr as f
    kw = dict(base_lr=5e-4, warmup_epochs=2, max_epochs=100, warmup_start_lr=1e-7, eta_min=1e-8)
    assert f(0, **kw) == 1e-7 and f(1, **kw) == pytest.approx(5e-4) and f(2, **kw) == pytest.approx(5e-4)
    assert f(51, **kw) == pytest.approx(1e-8 + 0.5 * (5e-4 - 1e-8) * (1 + math.cos(math.pi * 49 / 98)))
    assert f(100, **kw) == pytest.approx(1e-8)


@pytest.mark.skipif(not os.path.exists(REF_SCHED), reason="/root/reference not present on this box")
def test_lr_schedule_matches_reference_class():
    """the reference steps its chainable scheduler once per epoch (intermediate_downscaling.py:756)."""
    from orbit2_b200.trainer import warmup_cosine_lr as f
    spec = importlib.util.spec_from_file_location("ref_lr_scheduler", REF_SCHED)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=5e-4)
    kw = dict(warmup_epochs=5, max_epochs=40, warmup_start_lr=1e-7, eta_min=1e-8)
    sch = mod.LinearWarmupCosineAnnealingLR(opt, **kw)
    for epoch in range(40):
        assert opt.param_groups[0]["lr"] == pytest.approx(f(epoch, 5e-4, **kw), rel=1e-9, abs=1e-15), epoch
        opt.step()
        sch.step()


@pytest.mark.gpu
def test_short_run_and_checkpoint_roundtrip(tmp_path):
    from orbit2_b200 import trainer
    conf = trainer.load_config(os.path.join(ROOT, "configs", "interm_8m.yaml"))
    conf["trainer"]["batch_size"] = 2
    conf["model"].update(embed_dim=128, depth=2, num_heads=2, decoder_depth=2)
    logs = []
    hist, eng = trainer.train(conf, "ERA5_1", (8, 16), epochs=2, steps_per_epoch=3, ckpt_dir=str(tmp_path),
                              log=lambda *a, **k: logs.append(a))
    assert len(hist) == 2 and all(math.isfinite(h) for h in hist) and len(logs) == 2
    ck = torch.load(tmp_path / "interm_epoch_1.ckpt", weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict"} and ck["epoch"] == 1
    # torch.optim.AdamW accepts the optimizer state as is
    params = [torch.nn.Parameter(v.clone()) for v in ck["model_state_dict"].values()]
    opt = torch.optim.AdamW(params, lr=1.0)
    opt.load_state_dict(ck["optimizer_state_dict"])
    assert opt.param_groups[0]["betas"] == (0.9, 0.99)
    # resume: a fresh engine restored from the checkpoint continues exactly like the original one
    model2, _, eng2 = trainer.build(conf, "ERA5_1", (8, 16), torch.device("cuda"))
    assert trainer.load_checkpoint(str(tmp_path / "interm_epoch_1.ckpt"), eng2) == 2
    assert eng2.step_count == eng.step_count == 6
    assert torch.equal(eng2.flat_p, eng.flat_p) and torch.equal(eng2.flat_m, eng.flat_m) and torch.equal(eng2.flat_v, eng.flat_v)
    x, y, iv, ov = next(trainer.synthetic_loader(conf, "ERA5_1", (8, 16), 2, 1, torch.device("cuda"), 77))
    eng.lr = eng2.lr = 1e-4
    torch.manual_seed(5)                      # the YAML trains with dropout 0.1: same seed -> same masks
    a = eng.step(x, y)
    torch.manual_seed(5)
    b = eng2.step(x, y)
    assert torch.allclose(a, b, rtol=1e-4) and torch.allclose(eng.flat_p, eng2.flat_p, rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["replicated", "full_shard"])
def test_module_forward_sees_engine_updates(mode):
    """The engine's fused AdamW updates the parameters through raw pointers (no ``_version`` bump): the module's own
    forward (validation, tiled inference) must still run on the CURRENT weights.  Regression test for a stale bf16 operand
    cache: module forward after further engine steps == a fresh module loaded from ``full_state_dict()``."""
    from oracle import cases, reslim_oracle as O
    from orbit2_b200 import engine, losses
    from orbit2_b200.reslim import Res_Slim_ViT
    cfg = cases.get_case("tiny")

    def make():
        m = Res_Slim_ViT(cfg["default_vars"], cfg["init_img_size"], len(cfg["default_vars"]), cfg["out_channels"], 1,
                         patch_size=cfg["patch_size"], drop_path=0.0, drop_rate=0.0, learn_pos_emb=True,
                         embed_dim=cfg["embed_dim"], depth=cfg["depth"], decoder_depth=cfg["decoder_depth"],
                         num_heads=cfg["num_heads"], compute_dtype=torch.bfloat16)
        m.spatial_resolution = cfg["spatial_resolution"]
        return m
    model = make()
    model.load_state_dict(O.init_state_dict(cfg, seed=3))
    model = model.cuda()
    meta = losses.MetricsMetaInfo(cfg["in_vars"], cfg["out_vars"], None, None)
    eng = engine.TrainEngine(model, losses.METRICS_REGISTRY["mse"](aggregate_only=True, metainfo=meta), cfg["in_vars"],
                             cfg["out_vars"], cfg["var_weights"], lr=1e-2, shard_params=(mode == "full_shard"))
    x, y = O.synthetic_batch(cfg, 2, cfg["in_vars"], cfg["out_vars"], seed=3)
    x, y = x.cuda(), y.cuda()
    eng.step(x, y)
    model.eval()
    with torch.no_grad():
        first = model(x, cfg["in_vars"], cfg["out_vars"]).float().clone()       # a validation pass in between
    model.train()
    for _ in range(3):
        eng.step(x, y)
    model.eval()
    with torch.no_grad():
        got = model(x, cfg["in_vars"], cfg["out_vars"]).float()
    fresh = make()
    fresh.load_state_dict(eng.full_state_dict())
    fresh = fresh.cuda().eval()
    with torch.no_grad():
        want = fresh(x, cfg["in_vars"], cfg["out_vars"]).float()
    assert (want - first).abs().max() > 1e-3                  # the weights did move (lr 1e-2)
    assert torch.equal(got, want)


@pytest.mark.gpu
def test_npz_shards_training_and_validation(tmp_path):
    """The reference's shard layout end to end: raw npz -> GPU normalisation -> training steps -> validation metrics."""
    from orbit2_b200 import trainer
    from tests.test_data import IN_VARS, OUT_VARS, make_shards
    inp, out = make_shards(str(tmp_path), n_files=2, n_t=4, h=8, w=16)
    conf = trainer.load_config(os.path.join(ROOT, "configs", "interm_8m.yaml"))
    conf["trainer"].update(batch_size=2, buffer_size=3)
    conf["model"].update(embed_dim=128, depth=2, num_heads=2, decoder_depth=1)
    conf["data"]["low_res_dir"] = {"ERA5_1": inp}
    conf["data"]["high_res_dir"] = {"ERA5_1": out}
    conf["data"]["dict_in_variables"]["ERA5_1"] = IN_VARS
    conf["data"]["dict_out_variables"]["ERA5_1"] = OUT_VARS
    dm = trainer.npz_data(conf, "ERA5_1", torch.device("cuda"), 0, 1)
    grid = dm.get_data_dims()[0][2:]
    assert tuple(grid) == (8, 16)
    logs = []
    hist, eng = trainer.train(conf, "ERA5_1", tuple(grid), epochs=2, steps_per_epoch=3, log=lambda *a, **k: logs.append(a[0]),
                              data_module=dm)
    assert len(hist) == 2 and all(math.isfinite(h) for h in hist)
    assert all("val rmse" in l and "pearson" in l and "mean_bias" in l for l in logs)
    val = trainer.validate(eng, dm)
    assert set(val) == {"rmse", "pearson", "mean_bias"} and all(math.isfinite(v) for v in val.values())
    assert -1.0 <= val["pearson"] <= 1.0 and val["rmse"] > 0


def test_pretrain_loader_semantics():
    """trainer.load_pretrained_weights follows examples/intermediate_downscaling.py:116-153: unknown keys dropped, shape
    mismatches dropped except pos_embed, which is resampled bicubically to the model's token grid
    (components/pos_embed.py:73-99); everything else copied in place."""
    import torch.nn.functional as F
    from oracle import cases
    from orbit2_b200 import trainer
    from orbit2_b200.reslim import Res_Slim_ViT
    cfg = cases.get_case("tiny")

    def make(img, out_ch):
        return Res_Slim_ViT(cfg["default_vars"], img, len(cfg["default_vars"]), out_ch, 1, superres_mag=cfg["superres_mag"],
                            cnn_ratio=cfg["cnn_ratio"], patch_size=cfg["patch_size"], drop_path=0.0, drop_rate=0.0,
                            learn_pos_emb=True, embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                            decoder_depth=cfg["decoder_depth"], num_heads=cfg["num_heads"], mlp_ratio=cfg["mlp_ratio"])
    torch.manual_seed(0)
    src = make((16, 32), 2)                      # pretrained on a coarser grid, with a different number of output channels
    torch.manual_seed(1)
    dst = make((32, 64), 3)
    sd = {k: v.detach().clone() for k, v in src.state_dict().items()}
    sd["pos_embed"] = torch.randn_like(sd["pos_embed"])
    sd["not_in_the_model.weight"] = torch.zeros(3)
    before = {k: v.detach().clone() for k, v in dst.state_dict().items()}
    rep = trainer.load_pretrained_weights(dst, sd, log=lambda *_: None)
    after = dst.state_dict()
    p = cfg["patch_size"]
    oh, ow = 16 // p, 32 // p
    want = F.interpolate(sd["pos_embed"].reshape(1, oh, ow, -1).permute(0, 3, 1, 2), size=(32 // p, 64 // p), mode="bicubic",
                         align_corners=False).permute(0, 2, 3, 1).flatten(1, 2)
    assert torch.equal(after["pos_embed"], want)
    assert "not_in_the_model.weight" in rep["dropped"]
    mism = [k for k in sd if k in before and sd[k].shape != before[k].shape and k != "pos_embed"]
    assert mism and all(k in rep["dropped"] and torch.equal(after[k], before[k]) for k in mism)   # head.8 / path2.3 / conv_out
    same = [k for k in sd if k in before and sd[k].shape == before[k].shape]
    assert same and all(torch.equal(after[k], sd[k]) for k in same)
    assert set(rep["loaded"]) == set(same) | {"pos_embed"}


@pytest.mark.reference
def test_pos_embed_interpolation_matches_live_reference(monkeypatch):
    from oracle import ref_shim
    from orbit2_b200 import trainer
    ref_shim.load_reference()
    import importlib
    pe = importlib.import_module("climate_learn.models.hub.components.pos_embed")
    monkeypatch.setattr(torch.distributed, "get_rank", lambda: 1)
    ck = {"pos_embed": torch.randn(1, 8 * 16, 24)}

    class M:
        patch_size = 2
    want = dict(ck)
    pe.interpolate_pos_embed(M(), want, new_size=(24, 48))
    got = trainer.interpolate_pos_embed(ck["pos_embed"], 2, (24, 48))
    assert torch.equal(got, want["pos_embed"])
