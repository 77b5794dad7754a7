"""CPU: host-side logic of the drop-in module against the oracle -- the exact front-end collapse tables, the
position/resolution embedding, the state-dict ABI, clip specification, registry and flat-layout helpers."""
import numpy as np
import pytest
import torch

from tests.util import build_model, load_golden, rel


def collapse_reference(m, x, var_ids):
    """What o2_frontend_fwd + var_agg.proj compute, written in torch from the tables (float64)."""
    tab_s, tab_v = m.frontend_tables(var_ids)
    tab_s, tab_v = tab_s.double(), tab_v.double()
    B, V, H, W = x.shape
    p = m.patch_size
    gh, gw = H // p, W // p
    P = x.double().reshape(B, V, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, V, p * p)
    P1 = torch.cat([P, torch.ones_like(P[..., :1])], -1)
    a = torch.einsum("tvk,vhk->tvh", P1, tab_s).softmax(1)
    heads = tab_s.shape[1]
    coef = torch.einsum("tvh,tvk->thvk", a, P1).reshape(-1, heads, V * (p * p + 1))
    o = torch.einsum("thk,hke->the", coef, tab_v).reshape(B, gh * gw, -1)
    return torch.nn.functional.linear(o, m.var_agg.proj.weight.double(), m.var_agg.proj.bias.double())


@pytest.mark.parametrize("fixture", ["tiny_mse", "tiny_prism_mae_lat"])
def test_frontend_collapse_matches_oracle(fixture):
    from oracle import cases, reslim_oracle as O
    z, meta, sd, _ = load_golden(fixture)
    cfg = cases.get_case(meta[0])
    m = build_model(cfg, sd).double()
    x = torch.from_numpy(z["x"]).double()
    taps = {}
    O.forward({k: v.double() for k, v in sd.items()}, cfg, x, cfg["in_vars"], cfg["out_vars"], taps)
    ours = collapse_reference(m, x, m.get_var_ids(cfg["in_vars"]))
    assert rel(ours, taps["agg"]) < 1e-6             # tables are built in fp32
    gh, gw = cfg["img_size"][0] // 2, cfg["img_size"][1] // 2
    tok0 = ours + m.pos_res_embed(gh, gw, torch.float64).double()[None]
    assert rel(tok0, taps["tokens0"]) < 1e-6          # pos_res_embed is computed in fp32


def test_state_dict_abi():
    from oracle import cases, reslim_oracle as O
    for name in ("tiny", "8m"):
        cfg = cases.get_case(name)
        m = build_model(cfg)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert shapes == O.param_shapes(cfg)
        assert list(shapes) == list(O.param_shapes(cfg))      # registration order too


def test_init_like_reference():
    from oracle import cases, reslim_oracle as O
    cfg = cases.get_case("tiny")
    m = build_model(cfg)
    assert torch.count_nonzero(m.var_embed) == 0 and torch.count_nonzero(m.var_query) == 0     # res_slimvit.py:70,74
    pe = torch.from_numpy(O.sincos_2d(cfg["embed_dim"], 4, 8)).float()
    assert torch.allclose(m.pos_embed[0], pe)
    assert float(m.blocks[0].attn.qkv.weight.std()) == pytest.approx(0.02, rel=0.15)
    assert torch.count_nonzero(m.blocks[0].attn.qkv.bias) == 0
    assert torch.all(m.norm.weight == 1)


def test_clip_spec_and_registry():
    from orbit2_b200 import losses
    assert losses.clip_spec(["total_precipitation_24hr", "orography", "2m_temperature_max"]) == (0, 0b010)
    assert losses.clip_spec(["2m_temperature", "total_precipitation_24hr"]) == (1, 0)
    with pytest.raises(ValueError):
        losses.clip_spec(["2m_temperature"])                    # reference: list.index raises
    for k in ("mse", "mae", "lat_mse", "bayesian_tv"):
        assert k in losses.METRICS_REGISTRY and losses.METRICS_REGISTRY[k].name == k
    meta = losses.MetricsMetaInfo([], [], np.array([60.0, 0.0, -60.0]), None)
    lw = losses.LatWeightedMSE(True, meta).lat_weights
    w = np.cos(np.deg2rad([60.0, 0.0, -60.0]))
    assert lw.shape == (1, 1, 3, 1) and lw.dtype == torch.float64
    assert np.allclose(lw.reshape(-1).numpy(), w / w.mean())


def test_geometry_errors():
    from oracle import cases
    cfg = cases.get_case("tiny")
    m = build_model(cfg)
    x = torch.zeros(1, 8, 8, 16)
    g = m.geometry(x, cfg["in_vars"], cfg["out_vars"], torch.float32)
    assert (g.gh, g.gw, g.L, g.T, g.hd) == (4, 8, 32, 32, 64)
    assert g.idx7 == [cfg["in_vars"].index(v) for v in cfg["out_vars"] + ["land_sea_mask", "orography", "lattitude", "landcover"]]
    m.img_size = (9, 16)                                   # odd rows: the reference's unpatchify raises (181-row case)
    with pytest.raises(RuntimeError):
        m.geometry(torch.zeros(1, 8, 9, 16), cfg["in_vars"], cfg["out_vars"], torch.float32)
    m.img_size = (8, 16)
    with pytest.raises(ValueError):
        m.geometry(torch.zeros(1, 8, 10, 16), cfg["in_vars"], cfg["out_vars"], torch.float32)


def test_flat_layout_runs():
    from orbit2_b200.dp import FlatLayout
    lay = FlatLayout(["a", "b", "c", "d"], [5, 16, 3, 8])
    assert lay.range["b"] == (8, 24) and lay.padded["a"] == (0, 8) and lay.total == 40
    assert lay.runs(["a", "b"]) == [[0, 24]]
    assert lay.runs(["d", "a", "b"]) == [[0, 24], [32, 40]]
    assert lay.runs(["c"]) == [[24, 32]]


def test_attn_keep_mask_restatement():
    """oracle/dropout_mask.py restates the bit-sliced attention keep mask of csrc/common.cuh: checked here against a
    scalar Python evaluation of the documented formula, plus its keep rate and the bit <-> key bijection."""
    import math
    from oracle import dropout_mask as DM
    M32 = 0xFFFFFFFF

    def lb(x):
        x &= M32; x ^= x >> 16; x = (x * 0x21F0AAAD) & M32; x ^= x >> 15; x = (x * 0x735A2D97) & M32; x ^= x >> 15
        return x

    seed, site, N, heads = 0x1234_5678_9ABC_DEF0, 17, 70, 2
    for p in (0.1, 0.35):
        thr16 = math.floor(p * 65536)
        hi8, frac8 = thr16 >> 8, thr16 & 0xFF
        m = DM.attn_scaled_mask(seed, site, 1, heads, N, p)
        nkb = (N + 31) // 32
        key = lb(DM.site_key(seed, site) ^ ((1 * 0x9E3779B1) & M32))          # (b, h) = (0, 1)
        for q, k in ((0, 0), (3, 31), (5, 32), (69, 69), (17, 40), (64, 2)):
            base = lb(((q * nkb + (k >> 5)) & M32) ^ key)
            thr, ge = hi8 + (1 if (lb(base ^ 0x68E31DA4) >> 24) < frac8 else 0), M32      # dithered byte threshold
            for i, mul in enumerate(DM._KEEP_MUL):
                w = ((base * mul) & M32) ^ ((base * mul) >> 32)
                ge = (w & ge) if (thr >> i) & 1 else (w | ge)
            kk = k & 31
            bit = 7 - (kk >> 2) + 8 * (kk & 1) + 16 * ((kk >> 1) & 1)
            want = ((ge >> bit) & 1) * 65536.0 / (65536.0 - thr16)
            assert float(m[0, 1, q, k]) == want
    assert sorted(int(b) for b in DM.attn_keep_bit(torch.arange(32))) == list(range(32))
    big = DM.attn_scaled_mask(seed, site, 1, 1, 512, 0.1)
    assert abs(float((big > 0).double().mean()) - (1 - math.floor(0.1 * 65536) / 65536)) < 2e-3   # 0.1 to 16 bits, not 25 / 256
    # the exact byte decision U >= thr: the keep word must equal a direct comparison of the reconstructed bytes
    nkb = 512 // 32
    key0 = lb(DM.site_key(seed, site))
    for q, kb in ((0, 0), (100, 7), (511, 15)):
        base = lb(((q * nkb + kb) & M32) ^ key0)
        planes = [(((base * mul) & M32) ^ ((base * mul) >> 32)) for mul in DM._KEEP_MUL]
        thr = (math.floor(0.1 * 65536) >> 8) + (1 if (lb(base ^ 0x68E31DA4) >> 24) < (math.floor(0.1 * 65536) & 0xFF) else 0)
        for kk in range(32):
            bit = 7 - (kk >> 2) + 8 * (kk & 1) + 16 * ((kk >> 1) & 1)
            U = sum(((planes[i] >> bit) & 1) << i for i in range(8))
            assert (float(big[0, 0, q, 32 * kb + kk]) > 0) == (U >= thr)
    assert abs(float(big.mean()) - 1.0) < 4e-3                                  # E[mask] = 1: kept values carry 1 / keep_prob


def test_grad_scaler_state_machine():
    """engine.GradScaler follows torch's GradScaler update rule with the reference's settings (init 8192, growth interval
    100, floor 128: examples/intermediate_downscaling.py:493-495, 741-742); checked against torch.amp.GradScaler's own
    scale sequence for the same found-inf pattern."""
    from orbit2_b200.engine import GradScaler
    s = GradScaler()
    assert s.scale == 8192.0
    pattern = [False] * 99 + [True] + [False] * 100 + [False] * 100 + [True] * 9
    want = 8192.0
    tracker = 0
    for bad in pattern:
        s.update(bad)
        if bad:
            want, tracker = want * 0.5, 0
        else:
            tracker += 1
            if tracker == 100:
                want, tracker = want * 2.0, 0
        want = max(want, 128.0)
        assert s.scale == want
    assert s.scale == 128.0 and s.skipped == 10          # 4096 -> 8192 -> 16384, then nine halvings stop at the floor
    t = GradScaler(); t.load_state_dict(s.state_dict())
    assert (t.scale, t.growth_tracker) == (s.scale, s.growth_tracker)


def test_keep_masks_known_answer(golden_dir):
    """the dropout mask restatements (oracle/dropout_mask.py) against the committed known-answer bits
    (tests/golden/keep_masks.npz, written by oracle/make_mask_golden.py)."""
    import os
    from oracle import make_mask_golden as G
    z = np.load(os.path.join(golden_dir, "keep_masks.npz"))
    got = G.build()
    assert sorted(z.files) == sorted(got)
    for k in z.files:
        assert np.array_equal(z[k], got[k]), k


def test_drop_plan_redraw_keeps_buffers():
    """CUDA-graph mode keeps ONE DropPlan whose per-sample drop-path factors live in fixed tensors: redraw_paths must write
    new bernoulli(keep) / keep factors into the SAME storage (captured kernels keep reading those addresses), reproducibly
    for a given seed, and leave the blocks without drop-path (rate 0) alone."""
    from orbit2_b200.reslim import DropPlan
    dpr = [0.0, 0.1, 0.2, 0.3]
    plan = DropPlan(0.1, dpr, 64, seed=5, device="cpu")
    assert plan.path[0] == (None, None) and plan.branch_active(0)           # element dropout alone keeps the branch active
    ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in plan.path[1:]]
    before = [a.clone() for a, _ in plan.path[1:]]
    plan.redraw_paths(1234)
    assert [(a.data_ptr(), b.data_ptr()) for a, b in plan.path[1:]] == ptrs
    assert any(not torch.equal(x, a) for x, (a, _) in zip(before, plan.path[1:]))
    again = DropPlan(0.1, dpr, 64, seed=5, device="cpu")
    again.redraw_paths(1234)
    for (a, b), (c, d), dp in zip(plan.path[1:], again.path[1:], dpr[1:]):
        assert torch.equal(a, c) and torch.equal(b, d)
        for v in set(a.tolist()) | set(b.tolist()):
            assert v == 0.0 or abs(v - 1.0 / (1.0 - dp)) < 1e-6
    assert not DropPlan(0.0, [0.0, 0.0], 4, 1, "cpu").branch_active(1)
