"""GPU parity at BASELINE.json's FULL sizes (interm_117m: 180x360 -> 720x1440, L = 16200 tokens, 16 heads x 64, D = 1024).

The float64 oracle of the whole model is far too slow here, so these tests use (a) size-independent properties of the
path (softmax rows sum to one, batch independence, gradient sum rules, linearity of the loss gradient) and (b) float64
references that are cheap even at full size because they only follow a SAMPLE of rows / a single layer (a few
attention query rows against all 16200 keys, the head tail and the loss on the full 720x1440 grid)."""
import numpy as np
import pytest
import torch

from tests.util import rel

pytestmark = pytest.mark.gpu

N117, HEADS, HD = 16200, 16, 64


def _qkv(B, N, heads, hd, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(B * N, 3 * heads * hd, generator=g, device="cuda") * scale).to(torch.bfloat16)


def test_attention_full_length_sampled_rows():
    """L = 16200 (ragged: 16200 % 128 = 72), 16 heads: forward rows and dQ rows of 48 sampled queries, dK / dV rows of 48
    sampled keys against float64 on the same bf16 operands; the sampled-key check needs every query's softmax column,
    which is rebuilt in float64 from the kernel-independent float64 log-sum-exp, one head at a time."""
    from orbit2_b200 import ops
    B, N, heads, hd = 1, N117, HEADS, HD
    D = heads * hd
    qkv = _qkv(B, N, heads, hd, 1)
    g = torch.Generator(device="cuda").manual_seed(2)
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, B, N, heads, hd)
    dqkv = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd)
    t = qkv.double().reshape(N, 3, heads, hd)
    do = dout.double().reshape(N, heads, hd)
    o_k = out.double().reshape(N, heads, hd)
    dq_k, dk_k, dv_k = dqkv.double().reshape(N, 3, heads, hd).unbind(1)
    rows = torch.cat([torch.tensor([0, 1, 127, 128, N - 73, N - 72, N - 1], device="cuda"),
                      torch.randint(0, N, (41,), generator=g, device="cuda")])
    sc = hd ** -0.5
    worst = dict(o=0.0, lse=0.0, dq=0.0, dk=0.0, dv=0.0)
    for h in (0, 7, 15):
        q, k, v = t[:, 0, h], t[:, 1, h], t[:, 2, h]
        s = (q * sc) @ k.T                                   # [N, N] float64: 2.1 GB
        lse_ref = torch.logsumexp(s, -1)
        p = torch.exp(s - lse_ref[:, None])
        del s
        o_ref = p[rows] @ v
        worst["o"] = max(worst["o"], rel(o_k[rows, h], o_ref))
        worst["lse"] = max(worst["lse"], rel(lse[0, h], lse_ref))
        dp = do[:, h] @ v.T                                  # [N, N]
        delta = (do[:, h] * (p @ v)).sum(-1)
        ds = p * (dp - delta[:, None])
        del dp
        worst["dq"] = max(worst["dq"], rel(dq_k[rows, h], (ds[rows] @ k) * sc))
        worst["dk"] = max(worst["dk"], rel(dk_k[rows, h], (ds[:, rows].T @ q) * sc))
        worst["dv"] = max(worst["dv"], rel(dv_k[rows, h], p[:, rows].T @ do[:, h]))
        del p, ds
    assert worst["lse"] < 1e-3 and worst["o"] < 1.5e-2, worst
    assert worst["dq"] < 2e-2 and worst["dk"] < 2e-2 and worst["dv"] < 2e-2, worst


def test_attention_full_length_sum_rules():
    """Properties that hold for any length: with V = 1 every output is 1 (softmax rows sum to one), sum_k dV[k] = sum_q dO[q]
    (same reason), dQ = dK = 0 when V is constant (dP - delta vanishes), and every (batch, head) is independent of the
    others (B = 2 equals two B = 1 launches bit for bit)."""
    from orbit2_b200 import ops
    B, N, heads, hd = 2, N117, HEADS, HD
    D = heads * hd
    qkv = _qkv(B, N, heads, hd, 3)
    g = torch.Generator(device="cuda").manual_seed(4)
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, B, N, heads, hd)
    dqkv = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd)
    for b in range(B):
        o1, l1 = ops.attn_fwd(qkv[b * N:(b + 1) * N].contiguous(), 1, N, heads, hd)
        assert torch.equal(o1, out[b * N:(b + 1) * N]) and torch.equal(l1[0], lse[b])
        d1 = ops.attn_bwd(qkv[b * N:(b + 1) * N].contiguous(), o1, dout[b * N:(b + 1) * N].contiguous(), l1, 1, N, heads, hd)
        # dK / dV: single-owner sums, bit for bit; dQ: fp32 partials reduced in the L2 in a run-dependent order, equal to
        # fp32 rounding (a different bf16 neighbour is possible where the fp32 sum sits on a rounding boundary)
        full = dqkv[b * N:(b + 1) * N].view(N, 3, D)
        assert torch.equal(d1.view(N, 3, D)[:, 1:], full[:, 1:])
        assert rel(d1.view(N, 3, D)[:, 0], full[:, 0]) < 8e-3
        d2 = ops.attn_bwd(qkv[b * N:(b + 1) * N].contiguous(), o1, dout[b * N:(b + 1) * N].contiguous(), l1, 1, N, heads, hd,
                          two_pass=True)
        d3 = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd, two_pass=True)
        assert torch.equal(d2, d3[b * N:(b + 1) * N])             # the deterministic path: bit for bit
    dv = dqkv.double().reshape(B, N, 3, heads, hd)[:, :, 2]
    assert rel(dv.sum(1), dout.double().reshape(B, N, heads, hd).sum(1)) < 5e-3
    ones = qkv.clone().reshape(B * N, 3, D)
    ones[:, 2] = 1.0
    ones = ones.reshape(B * N, 3 * D)
    out1, lse1 = ops.attn_fwd(ones, B, N, heads, hd)
    assert (out1.float() - 1.0).abs().max().item() < 8e-3            # bf16 rounding of the normalised accumulator
    assert torch.equal(lse1, lse)
    d1 = ops.attn_bwd(ones, out1, dout, lse1, B, N, heads, hd).double().reshape(B * N, 3, D)
    scale = dqkv.double().abs().max().item()
    assert d1[:, 0].abs().max().item() < 2e-2 * scale and d1[:, 1].abs().max().item() < 2e-2 * scale


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_head_tail_and_loss_full_grid(dtype):
    """Residual branch + head tail (unpatchify index map, conv_out, GELU / PixelShuffle / conv, crop-add) and the fused
    clip + Bayesian-TV loss with latitude and variable weights on the full 720x1440 grid (B = 2), forward and backward,
    against the oracle's float64 functions run on the GPU."""
    from oracle import reslim_oracle as O
    from orbit2_b200 import _lib as L, ops
    B, C, gh, gw, p, mag, cr, V = 2, 3, 90, 180, 2, 4, 4, 23
    Hx, Wx, Ho, Wo = gh * p, gw * p, gh * p * mag, gw * p * mag
    g = torch.Generator(device="cuda").manual_seed(5)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
    x = rn(B, V, Hx, Wx)
    idx = [22, 5, 6, 0, 1, 2, 3]
    w1, b1 = rn(cr * mag * mag, C + 4, 3, 3) * 0.2, rn(cr * mag * mag) * 0.1
    w2, b2, wo, bo = rn(C, cr, 3, 3) * 0.3, rn(C) * 0.1, rn(C, C, 3, 3) * 0.3, rn(C) * 0.1
    ho = rn(B * gh * gw, C * (mag * p) ** 2).to(dtype)
    h1, g1 = ops.path2_conv1_fwd(x, idx, w1, b1, dtype, mag=mag)
    preds = ops.headtail_fwd(ho, h1, wo, bo, w2, b2, B, C, gh, gw, p, mag, g1=g1)
    d = lambda t: t.double().requires_grad_(True)
    w1d, b1d, w2d, b2d, wod, bod, hod = d(w1), d(b1), d(w2), d(b2), d(wo), d(bo), d(ho)
    sd = {"path2.0.weight": w1d, "path2.0.bias": b1d, "path2.3.weight": w2d, "path2.3.bias": b2d}
    p2 = O.path2(sd, x[:, idx].double(), mag)
    img = O.unpatchify(hod.reshape(B, gh * gw, -1), (Hx, Wx), p, mag, C)
    ref = torch.nn.functional.conv2d(img, wod, bod, padding=1) + p2[:, :, :Ho, :Wo]
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel(preds, ref.detach()) < tol
    # loss on OUR prediction (so both sides clip / difference the same field), gradient w.r.t. that prediction
    tgt = rn(B, C, Ho, Wo)
    tgt[:, 0].clamp_(min=0)
    lat = np.linspace(90, -90, Ho)
    lw = np.cos(np.deg2rad(lat))
    lw = torch.from_numpy(lw / lw.mean()).cuda()
    chw = torch.tensor([1.0, 10.0, 10.0], device="cuda")
    vec, dpred = ops.loss_fwd_bwd(preds, tgt, L.LOSS_BAYESIAN_TV, lat_w=lw.float(), ch_w=chw, clamp_ch=0)
    pr = preds.detach().double().requires_grad_(True)
    clipped = torch.cat([pr[:, :1].clamp(min=0), pr[:, 1:]], 1)
    names = ["total_precipitation_24hr", "2m_temperature_min", "2m_temperature_max"]
    lref = O.bayesian_tv(clipped, tgt.double(), names, {names[0]: 1.0, names[1]: 10.0, names[2]: 10.0}, False, lw.view(1, 1, Ho, 1))
    assert rel(vec, lref.detach()) < (1e-5 if dtype == torch.float32 else 1e-4)
    lref[-1].backward()
    assert rel(dpred, pr.grad) < (2e-5 if dtype == torch.float32 else 1e-2)
    # backward of the head tail with that gradient
    ref.backward(dpred.double())
    G = {k: torch.zeros_like(v) for k, v in dict(w1=w1, b1=b1, w2=w2, b2=b2, wo=wo, bo=bo).items()}
    dho, dh1 = ops.headtail_bwd(dpred, ho, h1, wo, w2, G["wo"], G["bo"], G["w2"], G["b2"], B, C, gh, gw, p, mag, g1=g1)
    ops.path2_conv1_bwd(x, idx, dh1, G["w1"], G["b1"])
    gt = 2e-5 if dtype == torch.float32 else 2e-2
    assert rel(dho, hod.grad) < gt
    for k, r in dict(wo=wod, bo=bod, w2=w2d, b2=b2d, w1=w1d, b1=b1d).items():
        assert rel(G[k], r.grad) < gt, k


def test_model_117m_batch_independence_and_grad_linearity():
    """interm_117m at its full 180x360 grid, bf16 arm: sample b of a B = 2 batch gives bit-identical predictions to the
    sample alone (no kernel mixes samples; tile schedules do not change per-row arithmetic), and the parameter gradients
    of the mean loss over the batch are the mean of the per-sample gradients (linearity of backward + reductions)."""
    from oracle import cases, reslim_oracle as O
    from orbit2_b200 import losses
    from tests.util import build_model
    cfg = cases.get_case("117m")
    torch.manual_seed(0)
    m = build_model(cfg, None, "cuda", torch.bfloat16)
    with torch.no_grad():
        m.var_embed.normal_(0, 0.02)
        m.var_query.normal_(0, 0.02)
    m.train()
    x, y = O.synthetic_batch(cfg, 2, cfg["in_vars"], cfg["out_vars"], seed=11)
    x, y = x.cuda(), y.cuda()
    meta = losses.MetricsMetaInfo(cfg["in_vars"], cfg["out_vars"], np.linspace(90, -90, 720), None)
    loss_fn = losses.METRICS_REGISTRY["lat_bayesian_tv"](aggregate_only=True, metainfo=meta)

    def run(xb, yb):
        for p_ in m.parameters():
            p_.grad = None
        pred = m(xb, cfg["in_vars"], cfg["out_vars"])
        loss = loss_fn(pred, yb, var_names=cfg["out_vars"], var_weights=cfg["var_weights"], clip_out_variables=cfg["out_vars"])
        loss.backward()
        return pred.detach().clone(), loss.item(), {k: p_.grad.detach().double().clone() for k, p_ in m.named_parameters()
                                                   if p_.grad is not None}

    pred2, loss2, g2 = run(x, y)
    assert pred2.shape == (2, 3, 720, 1440) and torch.isfinite(pred2.float()).all()
    parts = [run(x[b:b + 1], y[b:b + 1]) for b in range(2)]
    for b in range(2):
        assert torch.equal(parts[b][0][0], pred2[b]), f"sample {b} depends on its batch"
    assert abs(loss2 - 0.5 * (parts[0][1] + parts[1][1])) < 1e-4 * abs(loss2)
    bad = {}
    for k, gfull in g2.items():
        gsum = 0.5 * (parts[0][2][k] + parts[1][2][k])
        denom = gsum.abs().max().item()
        if denom == 0:
            continue
        e = (gfull - gsum).abs().max().item() / denom
        if e > 2e-2:
            bad[k] = e
    assert not bad, bad


def test_model_117m_whole_model_vs_oracle(golden_dir):
    """BASELINE configs[1] itself -- interm_117m on its full 180x360 -> 720x1440 grid (L = 16200 tokens, 126 M parameters),
    B = 1, the shipped Bayesian-TV training loss (res_slimvit.py:312-338 forward, intermediate_downscaling.py:281-306
    training_step) -- on the SAME seeded weights and batch as the oracle:

    (a) against the FLOAT64 oracle, stored compactly in tests/golden/117m_fullgrid_f64_compact.npz by
        oracle/make_golden_117m.py (the float64 run takes 4.5 min of host time: loss, 8192 sampled prediction elements,
        2048 sampled elements + the max-norm of EVERY parameter gradient): fp32 arm 1e-4, bf16 arm 2e-2 (north_star);
    (b) against the fp32 CPU oracle run live here (SDPA attention = the reference's FusedAttn.DEFAULT path), EVERY element
        of the prediction and of every parameter gradient.  That oracle is itself an fp32 computation: its own deviation
        from float64 is recorded per parameter in the fixture (o32_err: <= 3.6e-5 everywhere except pos_embed, 7.2e-4 --
        pos_embed's gradient is the raw, un-averaged token gradient after 8 blocks), so the bound for (b) is
        tol + 2 x o32_err[parameter]."""
    import os
    from oracle import cases, reslim_oracle as O
    from orbit2_b200 import losses
    from tests.util import build_model
    z = np.load(os.path.join(golden_dir, "117m_fullgrid_f64_compact.npz"))
    cfg = cases.get_case("117m")
    sd = O.init_state_dict(cfg, seed=0)
    x, y = O.synthetic_batch(cfg, 1, cfg["in_vars"], cfg["out_vars"], seed=0)
    chk = sum(v.double().abs().sum().item() for v in sd.values())
    assert abs(chk - float(z["sd_checksum"])) < 1e-9 * chk, "the seeded init differs from the one the fixture was made with"
    names = [k[len("g_absmax/"):] for k in z.files if k.startswith("g_absmax/")]
    assert len(names) >= 150                                 # 23 patch embeds, 8 blocks x 12, head, convs, embeddings
    # live fp32 CPU oracle (all elements)
    torch.set_num_threads(os.cpu_count() or 1)
    old = O.USE_SDPA
    O.USE_SDPA = True
    try:
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        taps = {}
        loss32 = O.training_step(sdr, cfg, x, y, cfg["in_vars"], cfg["out_vars"], "bayesian_tv", cfg["var_weights"], None, taps)
        loss32.backward()
    finally:
        O.USE_SDPA = old
    assert abs(loss32.item() - float(z["loss_f32_oracle"])) < 1e-5 * abs(loss32.item())   # same oracle, same inputs
    pred32 = taps["preds"].detach()
    g32 = {k: v.grad for k, v in sdr.items() if v.grad is not None}
    for dtype, tol in ((torch.float32, 1e-4), (torch.bfloat16, 2e-2)):
        m = build_model(cfg, sd, "cuda", dtype)
        m.train()
        meta = losses.MetricsMetaInfo(cfg["in_vars"], cfg["out_vars"], None, None)
        loss_fn = losses.METRICS_REGISTRY["bayesian_tv"](aggregate_only=False, metainfo=meta)
        pred = m(x.cuda(), cfg["in_vars"], cfg["out_vars"])
        vec = loss_fn(pred, y.cuda(), var_names=cfg["out_vars"], var_weights=cfg["var_weights"],
                      clip_out_variables=cfg["out_vars"])
        vec[-1].backward()
        torch.cuda.synchronize()
        pc = pred.detach().double().cpu().clone()         # the oracle's tap is the clipped prediction (channel 0 = precipitation)
        pc[:, 0].clamp_(min=0)
        grads = {k: p.grad.detach().double().cpu() for k, p in m.named_parameters() if p.grad is not None}
        # (a) float64 fixture
        e_loss = abs(vec[-1].item() - float(z["loss"])) / abs(float(z["loss"]))
        e_pred = np.abs(pc.reshape(-1).numpy()[z["pred_idx"]] - z["pred_val"]).max() / float(z["pred_absmax"])
        worst64 = {k: np.abs(grads[k].reshape(-1).numpy()[z["g_idx/" + k]] - z["g_val/" + k]).max() / float(z["g_absmax/" + k])
                   for k in names}
        top64 = sorted(worst64.items(), key=lambda kv: -kv[1])[:4]
        # (b) live fp32 oracle, every element
        e_pred32 = rel(pc, pred32)
        worst32 = {k: rel(grads[k], g32[k]) for k in names}
        over32 = {k: v for k, v in worst32.items() if v > tol + 2.0 * float(z["o32_err/" + k])}
        print(f"117m whole-model parity {dtype}: vs float64 fixture: loss {e_loss:.2e} pred {e_pred:.2e} grads {top64}; "
              f"vs live fp32 oracle: pred {e_pred32:.2e} worst grad {max(worst32.items(), key=lambda kv: kv[1])}")
        assert e_loss < tol and e_pred < tol, (dtype, e_loss, e_pred)
        bad = {k: v for k, v in worst64.items() if v > tol}
        assert not bad, (dtype, "vs float64", bad)
        assert e_pred32 < tol + 2.0 * float(z["pred_o32_err"]) and not over32, (dtype, "vs live fp32 oracle", e_pred32, over32)
        del m, pred, vec, grads
        torch.cuda.empty_cache()
