"""CPU tests: pin the oracle restatement (oracle/reslim_oracle.py) to the reference.

* against the committed golden fixtures (outputs + every parameter gradient of the LIVE reference,
  float64, written by oracle/make_golden.py) -- runs everywhere;
* against the live reference module itself when /root/reference is present (build container).
"""
import os

import numpy as np
import pytest
import torch

from oracle import cases, reslim_oracle as O

FIXTURES = [("tiny_mse.npz", "tiny", "mse", False), ("tiny_bayesian_tv_lat.npz", "tiny", "bayesian_tv", True),
            ("tiny_prism_mae_lat.npz", "tiny_prism", "mae", True)]


def _load(golden_dir, fn):
    z = np.load(os.path.join(golden_dir, fn))
    sd = {k[2:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("w/")}
    gr = {k[2:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("g/")}
    return z, sd, gr


@pytest.mark.parametrize("fn,case,loss_name,use_lat", FIXTURES)
def test_oracle_matches_golden(golden_dir, fn, case, loss_name, use_lat):
    z, sd, gr = _load(golden_dir, fn)
    cfg = cases.get_case(case)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = torch.from_numpy(z["x"]).double()
    y = torch.from_numpy(z["y"]).double()
    lw = O.lat_weights(z["lat"]) if use_lat else None
    taps = {}
    loss = O.training_step(sd, cfg, x, y, cfg["in_vars"], cfg["out_vars"], loss_name, cfg["var_weights"], lw, taps)
    loss.backward()
    assert abs(loss.item() - z["loss_vec"][-1]) <= 1e-6 * abs(z["loss_vec"][-1])
    np.testing.assert_allclose(taps["preds"].detach().numpy(), z["pred"], rtol=1e-5, atol=1e-6)
    for k, g in gr.items():
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])
        scale = g.abs().max().item() + 1e-12
        # fixtures store gradients as float32 of float64 results
        assert (got - g).abs().max().item() <= 2e-6 * scale + 1e-12, k


def test_oracle_loss_vectors(golden_dir):
    z = np.load(os.path.join(golden_dir, "loss_vectors.npz"))
    pred0 = torch.from_numpy(z["pred"]); tgt = torch.from_numpy(z["target"])
    for use_lat in (False, True):
        sfx = "_lat" if use_lat else ""
        lw = O.lat_weights(z["lat"]) if use_lat else None
        for nm, fn in (("mse", O.mse), ("bayesian_tv", O.bayesian_tv)):
            p = pred0.clone().requires_grad_(True)
            v = fn(p, tgt, cases.OUT_VARS_3, cases.VAR_WEIGHTS, False, lw)
            v[-1].backward()
            np.testing.assert_allclose(v.detach().numpy(), z[nm + sfx], rtol=1e-12)
            np.testing.assert_allclose(p.grad.numpy(), z[nm + sfx + "_grad"], rtol=1e-10, atol=1e-15)
        p = pred0.clone().requires_grad_(True)
        v = O.mae(p, tgt, False, lw)
        v[-1].backward()
        np.testing.assert_allclose(v.detach().numpy(), z["mae" + sfx], rtol=1e-12)
        np.testing.assert_allclose(p.grad.numpy(), z["mae" + sfx + "_grad"], rtol=1e-10, atol=1e-15)


def test_unpatchify_index_map():
    """the flat reinterpretation, checked with an index tensor."""
    H, W, p, mag, C = 4, 8, 2, 4, 3
    L = (H // p) * (W // p)
    K = C * (mag * p) ** 2
    x = torch.arange(L * K, dtype=torch.float64).reshape(1, L, K)
    img = O.unpatchify(x, (H, W), p, mag, C)
    w = W * mag // p
    for l in range(L):
        for k in range(K):
            cell = l * (K // (p * p * C)) + k // (p * p * C)
            hh, ww = divmod(cell, w)
            r = k % (p * p * C)
            pp, qq, cc = r // (p * C), (r // C) % p, r % C
            assert img[0, cc, p * hh + pp, p * ww + qq].item() == x[0, l, k].item()


def test_reference_error_paths():
    cfg = cases.get_case("tiny")
    with pytest.raises(ValueError):
        O.find_var_index(["a", "b"], ["a"])                      # static field missing
    with pytest.raises(ValueError):
        O.clip_replace_constant(torch.zeros(1, 1, 2, 2), torch.zeros(1, 1, 2, 2), ["2m_temperature"])


@pytest.mark.reference
@pytest.mark.parametrize("case,B", [("tiny", 2), ("tiny_prism", 1), ("8m", 1), ("1b_small", 1), ("10b_small", 1)])
def test_oracle_matches_live_reference(case, B):
    from oracle import make_golden, ref_shim
    ref = ref_shim.load_reference()
    cfg = cases.get_case(case)
    sd = {k: v.double() for k, v in O.init_state_dict(cfg, 3).items()}
    x, y = O.synthetic_batch(cfg, B, cfg["in_vars"], cfg["out_vars"], 3)
    m = make_golden.build_reference_model(ref, cfg, sd, torch.float64)
    want = m.forward(x.double(), list(cfg["in_vars"]), list(cfg["out_vars"]))
    got = O.forward(sd, cfg, x.double(), cfg["in_vars"], cfg["out_vars"])
    assert (want - got).abs().max().item() < 1e-12


@pytest.mark.reference
def test_oracle_dropout_sites_match_live_reference(monkeypatch):
    """Training-mode forward of the LIVE reference (drop_rate 0.1, drop_path 0.2) with every nn.Dropout / DropPath draw
    recorded, replayed through the oracle's ``masks``: pins the position and scaling of all seven dropout sites
    (res_slimvit.py:284, attention.py:75,81, mlp.py:65,68, vit_blocks.py:78-79)."""
    from oracle import make_golden, ref_shim
    ref = ref_shim.load_reference()
    cfg = cases.get_case("tiny")
    sd = {k: v.double() for k, v in O.init_state_dict(cfg, 5).items()}
    x, _ = O.synthetic_batch(cfg, 2, cfg["in_vars"], cfg["out_vars"], 5)
    m = ref.Res_Slim_ViT(cfg["default_vars"], cfg["init_img_size"], len(cfg["default_vars"]), cfg["out_channels"],
                         history=1, superres_mag=cfg["superres_mag"], cnn_ratio=cfg["cnn_ratio"],
                         patch_size=cfg["patch_size"], drop_path=0.2, drop_rate=0.1, learn_pos_emb=True,
                         embed_dim=cfg["embed_dim"], depth=cfg["depth"], decoder_depth=cfg["decoder_depth"],
                         num_heads=cfg["num_heads"], mlp_ratio=cfg["mlp_ratio"], FusedAttn_option=ref.FusedAttn.NONE)
    m.load_state_dict(sd, strict=True)
    m = m.double().train()
    m.spatial_resolution = cfg["spatial_resolution"]
    rec = []

    def fake_dropout(inp, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return inp
        mask = (torch.rand(inp.shape, dtype=inp.dtype) >= p).to(inp.dtype) / (1.0 - p)
        rec.append(("drop", mask))
        return inp * mask

    def fake_droppath(self, t):
        if self.drop_prob == 0.0 or not self.training:
            return t
        keep = 1.0 - self.drop_prob
        mask = t.new_empty((t.shape[0],) + (1,) * (t.ndim - 1)).bernoulli_(keep) / keep
        rec.append(("path", mask))
        return t * mask

    monkeypatch.setattr(torch.nn.functional, "dropout", fake_dropout)
    monkeypatch.setattr(ref_shim._DropPath, "forward", fake_droppath)
    torch.manual_seed(11)
    want = m.forward(x.double(), list(cfg["in_vars"]), list(cfg["out_vars"]))
    monkeypatch.undo()
    # call order: pos_drop; per block attn_drop, proj_drop, [drop_path1], drop1, drop2, [drop_path2]
    # (dpr = linspace(0, 0.2, depth): block 0 has Identity instead of DropPath, res_slimvit.py:84, vit_blocks.py:61)
    it = iter(rec)
    masks = {"pos": next(it)[1]}
    for i in range(cfg["depth"]):
        b = f"blocks.{i}."
        has_path = i > 0
        order = ["attn", "proj"] + (["path1"] if has_path else []) + ["drop1", "drop2"] + (["path2"] if has_path else [])
        for key in order:
            kind, mk = next(it)
            assert kind == ("path" if key.startswith("path") else "drop"), (b, key, kind)
            masks[b + key] = mk
    assert next(it, None) is None
    got = O.forward(sd, cfg, x.double(), cfg["in_vars"], cfg["out_vars"], None, masks)
    assert (want - got).abs().max().item() < 1e-12
    plain = O.forward(sd, cfg, x.double(), cfg["in_vars"], cfg["out_vars"])
    assert (want - plain).abs().max().item() > 1e-3          # the masks do something


@pytest.mark.reference
def test_reference_bugs_documented():
    """SURVEY.md headline 4: odd H raises in the reference (unpatchify), 180 rows works."""
    from oracle import make_golden, ref_shim
    ref = ref_shim.load_reference()
    cfg = cases.get_case("tiny")
    cfg["img_size"] = (9, 16); cfg["init_img_size"] = (9, 16)
    sd = O.init_state_dict(cfg, 0)
    m = make_golden.build_reference_model(ref, cfg, sd, torch.float32)
    x, _ = O.synthetic_batch(cfg, 1, cfg["in_vars"], cfg["out_vars"], 0)
    with pytest.raises(RuntimeError):
        m.forward(x, list(cfg["in_vars"]), list(cfg["out_vars"]))


def test_oracle_matches_8m_golden(golden_dir):
    """BASELINE.json configs[0] (interm_8m, 23 variables, 32x64 -> 128x256) pinned to the LIVE reference: compact fixture
    written by oracle/make_golden.py (inputs, prediction, loss vector, gradients of a parameter subset that touches the
    front end, attention, MLP, head, residual branch and conv tail; weights = the seeded reference init, checksummed)."""
    from oracle import make_golden
    z = np.load(os.path.join(golden_dir, "8m_bayesian_tv_lat_compact.npz"))
    cfg = cases.get_case("8m")
    sd = {k: v.double() for k, v in O.init_state_dict(cfg, 0).items()}
    np.testing.assert_allclose(make_golden.weight_checksum(sd), z["w_checksum"], rtol=1e-6)   # fp32-stored weights
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    taps = {}
    loss = O.training_step(sd, cfg, torch.from_numpy(z["x"]).double(), torch.from_numpy(z["y"]).double(), cfg["in_vars"],
                           cfg["out_vars"], "bayesian_tv", cfg["var_weights"], O.lat_weights(z["lat"]), taps)
    loss.backward()
    assert abs(loss.item() - z["loss_vec"][-1]) <= 1e-6 * abs(z["loss_vec"][-1])
    np.testing.assert_allclose(taps["preds"].detach().numpy(), z["pred"], rtol=2e-5, atol=2e-6)
    n = 0
    for k in z.files:
        if not k.startswith("g"):
            continue
        name = k.split("/", 1)[1]
        got = sd[name].grad[:8] if k.startswith("g8/") else sd[name].grad
        want = torch.from_numpy(z[k]).double()
        assert (got - want).abs().max().item() <= 2e-6 * (want.abs().max().item() + 1e-12), k
        n += 1
    assert n == len(make_golden.COMPACT_GRADS)


@pytest.mark.reference
@pytest.mark.parametrize("case", ["8m", "1b_small", "10b_small"])
def test_oracle_gradients_match_live_reference(case):
    """Forward + clip + latitude-weighted bayesian_tv + backward of the LIVE reference (float64) against the restatement
    at the BASELINE widths the tiny fixtures do not cover: interm_8m (23 variables), interm_1b (head dim 128, PRISM
    7-variable batch), interm_10b (head dim 256).  Every parameter gradient."""
    from oracle import make_golden
    ref = make_golden.run_case(case, 1, "bayesian_tv", 3, True)
    cfg = cases.get_case(case)
    sd = {k[2:]: torch.from_numpy(v).double().requires_grad_(True) for k, v in ref.items() if k.startswith("w/")}
    loss = O.training_step(sd, cfg, torch.from_numpy(ref["x"]).double(), torch.from_numpy(ref["y"]).double(), cfg["in_vars"],
                           cfg["out_vars"], "bayesian_tv", cfg["var_weights"], O.lat_weights(ref["lat"]), {})
    loss.backward()
    # run_case stores weights / gradients as float32 of the float64 run: compare at that resolution
    assert abs(loss.item() - ref["loss_vec"][-1]) <= 2e-6 * abs(ref["loss_vec"][-1])
    for k, v in ref.items():
        if not k.startswith("g/"):
            continue
        g = sd[k[2:]].grad
        got = g if g is not None else torch.zeros_like(sd[k[2:]])
        want = torch.from_numpy(v).double()
        assert (got - want).abs().max().item() <= 5e-6 * (want.abs().max().item() + 1e-12), k
