"""north_star: 'loss curves matching over 100 steps'.  The training engine (kernels + fused AdamW, no autograd graph) is
run for 100 steps on the tiny config and compared step by step with the CPU oracle trained by torch.optim.AdamW
(same init, same batch sequence, dropout 0): fp32 arm within 1e-3 relative at every step, bf16 arm within 2e-2."""
import numpy as np
import pytest
import torch

from tests.util import build_model

pytestmark = pytest.mark.gpu


def oracle_curve(cfg, sd0, batches, steps, lr, betas, wd):
    from oracle import reslim_oracle as O
    sd = {k: v.clone().double().requires_grad_(True) for k, v in sd0.items()}
    opt = torch.optim.AdamW(list(sd.values()), lr=lr, betas=betas, weight_decay=wd)
    out = []
    for i in range(steps):
        x, y = batches[i % len(batches)]
        opt.zero_grad()
        loss = O.training_step(sd, cfg, x.double(), y.double(), cfg["in_vars"], cfg["out_vars"], "bayesian_tv",
                               cfg["var_weights"])
        loss.backward()
        opt.step()
        out.append(loss.item())
    return np.array(out)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 2e-2)])
def test_loss_curve_100_steps(dtype, tol):
    from oracle import cases, reslim_oracle as O
    from orbit2_b200 import engine, losses
    cfg = cases.get_case("tiny")
    steps, lr, betas, wd = 100, 1e-3, (0.9, 0.99), 1e-5
    sd0 = O.init_state_dict(cfg, seed=5)
    batches = [O.synthetic_batch(cfg, 2, cfg["in_vars"], cfg["out_vars"], seed=s) for s in range(4)]
    ref = oracle_curve(cfg, sd0, batches, steps, lr, betas, wd)
    m = build_model(cfg, sd0, "cuda", dtype)
    loss_fn = losses.METRICS_REGISTRY["bayesian_tv"](aggregate_only=True,
                                                      metainfo=losses.MetricsMetaInfo(cfg["in_vars"], cfg["out_vars"], None, None))
    eng = engine.TrainEngine(m, loss_fn, cfg["in_vars"], cfg["out_vars"], cfg["var_weights"], lr=lr, betas=betas,
                             weight_decay=wd)
    dev = [(x.cuda(), y.cuda()) for x, y in batches]
    ours = []
    for i in range(steps):
        x, y = dev[i % len(dev)]
        ours.append(eng.step(x, y)[-1])
    ours = torch.stack(ours).double().cpu().numpy()
    rel = np.abs(ours - ref) / np.abs(ref)
    assert ref[-1] < 0.95 * ref[0], "the oracle did not train: test is not meaningful"
    print("max rel deviation over 100 steps:", float(rel.max()), "first", ours[0], ref[0], "last", ours[-1], ref[-1])
    assert rel.max() < tol, (float(rel.max()), int(rel.argmax()), ours[:3], ref[:3], ours[-3:], ref[-3:])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cuda_graph_step_equals_eager(dtype):
    """TrainEngine.enable_graph: the step captured once as a CUDA graph (AdamW scalars read from device memory, batch
    copied into the static buffers) follows the eager engine step by step -- same kernels, same order -- also across a
    learning-rate change between replays."""
    from oracle import cases, reslim_oracle as O
    from orbit2_b200 import engine, losses
    cfg = cases.get_case("tiny")
    sd0 = O.init_state_dict(cfg, seed=7)
    batches = [tuple(t.cuda() for t in O.synthetic_batch(cfg, 2, cfg["in_vars"], cfg["out_vars"], seed=s)) for s in range(3)]
    curves, params = [], []
    for graph in (False, True):
        m = build_model(cfg, sd0, "cuda", dtype)
        loss_fn = losses.METRICS_REGISTRY["bayesian_tv"](
            aggregate_only=True, metainfo=losses.MetricsMetaInfo(cfg["in_vars"], cfg["out_vars"], None, None))
        eng = engine.TrainEngine(m, loss_fn, cfg["in_vars"], cfg["out_vars"], cfg["var_weights"], lr=1e-3,
                                 betas=(0.9, 0.99), weight_decay=1e-5)
        if graph:
            eng.enable_graph(warm_steps=2)
        out = []
        for i in range(12):
            if i == 8:
                eng.lr = 3e-4
            x, y = batches[i % 3]
            out.append(eng.step(x, y)[-1].clone())
        assert (eng._graph is not None) == graph
        curves.append(torch.stack(out).double().cpu())
        params.append(eng.flat_p.clone())
    tol = 1e-6 if dtype == torch.float32 else 1e-3           # same kernels; only atomics ordering may differ
    assert ((curves[0] - curves[1]).abs() / curves[0].abs()).max().item() < tol
    assert (params[0] - params[1]).abs().max().item() / params[0].abs().max().item() < tol
    with pytest.raises(RuntimeError):
        eng.step(batches[0][0][:1], batches[0][1][:1])       # batch shape is frozen into the graph


def test_cuda_graph_step_with_dropout():
    """Graph mode at the YAMLs' dropout 0.1 / drop-path 0.1 (interm_8m trains like that): the captured step draws NEW masks
    on every replay (device step word + refreshed drop-path factors) -- the same batch gives different losses step to step
    and with another torch seed, the same ones with the same seed -- and it trains."""
    from oracle import cases, reslim_oracle as O
    from orbit2_b200 import engine, losses
    cfg = cases.get_case("tiny")
    sd0 = O.init_state_dict(cfg, seed=7)
    x, y = (t.cuda() for t in O.synthetic_batch(cfg, 4, cfg["in_vars"], cfg["out_vars"], seed=0))

    def run(seed, lr):
        torch.manual_seed(seed)
        m = build_model(cfg, sd0, "cuda", torch.bfloat16, drop_rate=0.1, drop_path=0.1).train()
        loss_fn = losses.METRICS_REGISTRY["bayesian_tv"](
            aggregate_only=True, metainfo=losses.MetricsMetaInfo(cfg["in_vars"], cfg["out_vars"], None, None))
        eng = engine.TrainEngine(m, loss_fn, cfg["in_vars"], cfg["out_vars"], cfg["var_weights"], lr=lr, betas=(0.9, 0.99),
                                 weight_decay=1e-5)
        eng.enable_graph(warm_steps=2)
        out = [eng.step(x, y)[-1].item() for _ in range(40)]
        assert eng._graph is not None and eng._drop_word is not None
        return out
    frozen = run(1, 0.0)                                       # lr = 0: the weights never move, only the masks do
    replays = frozen[3:]
    assert len(set(replays)) > len(replays) // 2, "the replayed step keeps drawing the same masks"
    assert run(1, 0.0) == frozen                               # reproducible under torch.manual_seed
    assert run(2, 0.0)[3:] != replays
    trained = run(1, 2e-3)
    assert sum(trained[-5:]) < 0.95 * sum(trained[:5])


def test_grad_scaler_semantics():
    """bf16 branch of the reference driver (intermediate_downscaling.py:733-742): gradients are produced pre-scaled, the
    update un-scales them (same trajectory as the unscaled engine up to bf16 rounding of the scaled dL/dpred), and a step
    whose gradients contain inf / NaN is skipped, halving the scale."""
    from oracle import cases, reslim_oracle as O
    from orbit2_b200 import engine, losses, ops
    cfg = cases.get_case("tiny")
    sd0 = O.init_state_dict(cfg, seed=9)
    x, y = (t.cuda() for t in O.synthetic_batch(cfg, 2, cfg["in_vars"], cfg["out_vars"], seed=1))

    def make(scaler):
        m = build_model(cfg, sd0, "cuda", torch.bfloat16)
        loss_fn = losses.METRICS_REGISTRY["bayesian_tv"](
            aggregate_only=True, metainfo=losses.MetricsMetaInfo(cfg["in_vars"], cfg["out_vars"], None, None))
        return engine.TrainEngine(m, loss_fn, cfg["in_vars"], cfg["out_vars"], cfg["var_weights"], lr=1e-3,
                                  betas=(0.9, 0.99), weight_decay=1e-5, grad_scaler=scaler)
    plain, scaled = make(None), make(engine.GradScaler(growth_interval=3))
    for i in range(5):
        a, b = plain.step(x, y)[-1].item(), scaled.step(x, y)[-1].item()
        assert abs(a - b) / abs(a) < 2e-3, (i, a, b)                     # power-of-two scale: same numbers up to rounding
    assert scaled.scaler.scale == 8192.0 * 2 and scaled.step_count == 5  # one growth after three clean steps
    assert (plain.flat_p - scaled.flat_p).abs().max().item() / plain.flat_p.abs().max().item() < 2e-3
    # the non-finite test itself
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    g = torch.randn(100003, device="cuda")
    ops.nonfinite(g, flag); assert flag.item() == 0
    g[77777] = float("inf"); ops.nonfinite(g, flag); assert flag.item() == 1
    flag.zero_(); g[77777] = 0.0; g[100002] = float("nan"); ops.nonfinite(g[1:], flag); assert flag.item() == 1
    # an overflowing step is skipped: parameters and Adam state untouched, scale halved
    before, s0 = scaled.flat_p.clone(), scaled.scaler.scale
    bad_x = x.clone(); bad_x[0, 0, 0, 0] = float("inf")
    scaled.step(bad_x, y)
    assert scaled.step_count == 5 and scaled.scaler.scale == s0 / 2 and scaled.scaler.skipped == 1
    assert torch.equal(before, scaled.flat_p)
    scaled.step(x, y)
    assert scaled.step_count == 6 and not torch.equal(before, scaled.flat_p)
