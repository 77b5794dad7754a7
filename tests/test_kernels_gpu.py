"""GPU parity of the individual kernels (through the C ABI) against plain PyTorch / the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DT = [torch.float32, torch.bfloat16]


def rel(a, b):
    return (a.double() - b.double()).abs().max().item() / (b.double().abs().max().item() + 1e-12)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("T,D", [(37, 256), (1000, 1024), (129, 128), (64, 2048), (67, 3072), (35, 8192), (5003, 768),
                                 (3, 1024)])
def test_layernorm(dtype, T, D):
    """D = 3072 / 8192 are the interm_1b / 10b widths (rows shared by 2-8 warps)."""
    from orbit2_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(T + D)
    x = (torch.randn(T, D, generator=g, device="cuda") * 2 + 0.5).to(dtype)
    gamma = torch.randn(D, generator=g, device="cuda")
    beta = torch.randn(D, generator=g, device="cuda")
    dy = torch.randn(T, D, generator=g, device="cuda").to(dtype)
    dres = torch.randn(T, D, generator=g, device="cuda").to(dtype)
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    xr = x.double().requires_grad_(True); gr = gamma.double().requires_grad_(True); br = beta.double().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), gr, br, 1e-5)
    tol = 2e-6 if dtype == torch.float32 else 1e-2
    assert rel(y, yr.detach()) < tol
    yr.backward(dy.double())
    dgamma = torch.zeros(D, device="cuda"); dbeta = torch.zeros(D, device="cuda")
    dx = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dres=dres)
    assert rel(dx, xr.grad + dres.double()) < (1e-5 if dtype == torch.float32 else 1.5e-2)
    assert rel(dgamma, gr.grad) < (1e-5 if dtype == torch.float32 else 1e-2)
    assert rel(dbeta, br.grad) < (1e-5 if dtype == torch.float32 else 1e-2)
    dx2 = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dres=None)
    assert rel(dx2, xr.grad) < (1e-5 if dtype == torch.float32 else 1.5e-2)


@pytest.mark.parametrize("B,N,heads,hd", [(2, 200, 2, 64), (1, 64, 4, 32), (1, 333, 1, 64)])
def test_attn_simt(B, N, heads, hd):
    from orbit2_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N)
    D = heads * hd
    qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda")
    dout = torch.randn(B * N, D, generator=g, device="cuda")
    out, lse = ops.attn_fwd(qkv, B, N, heads, hd)
    t = qkv.double().reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4).requires_grad_(True)
    q, k, v = t.unbind(0)
    s = (q * hd ** -0.5) @ k.transpose(-2, -1)
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * N, D)
    assert rel(out, o.detach()) < 1e-5
    assert rel(lse, torch.logsumexp(s, -1).detach()) < 1e-5
    o.backward(dout.double())
    dqkv = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd)
    ref = t.grad.permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D)
    assert rel(dqkv, ref) < 2e-5


def test_elementwise():
    from orbit2_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(100003, generator=g, device="cuda")
    assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))
    for dtype in DT:
        X = torch.randn(777, 512, generator=g, device="cuda").to(dtype)
        out = torch.ones(512, device="cuda")
        ops.colsum(X, out)
        assert rel(out, X.double().sum(0) + 1) < 1e-5
    # AdamW vs torch.optim.AdamW over 3 steps
    p = torch.randn(5000, generator=g, device="cuda"); p0 = p.clone()
    m = torch.zeros_like(p); v = torch.zeros_like(p); pb = torch.empty(5000, device="cuda", dtype=torch.bfloat16)
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([pr], lr=5e-4, betas=(0.9, 0.99), weight_decay=1e-5)
    for step in range(1, 4):
        gr = torch.randn(5000, generator=g, device="cuda")
        ops.adamw(p, gr, m, v, pb, 5e-4, 0.9, 0.99, 1e-8, 1e-5, step)
        pr.grad = gr.clone(); opt.step()
    assert rel(p, pr.detach()) < 1e-6
    assert torch.equal(pb, p.to(torch.bfloat16))
    # vector body + scalar tail (n % 4 != 0) and a slice at a 4-byte offset (all-scalar path) give the same update
    n = 5003
    src = [torch.randn(n, generator=g, device="cuda") for _ in range(3)]          # p, g, v (made non-negative)
    res = []
    for off in (0, 1):
        p_, g_, m_, v_ = (torch.empty(n + off, device="cuda")[off:] for _ in range(4))
        p_.copy_(src[0]); g_.copy_(src[1]); m_.zero_(); v_.copy_(src[2].abs())
        ops.adamw(p_, g_, m_, v_, None, 5e-4, 0.9, 0.99, 1e-8, 1e-5, 2)
        res.append((p_.clone(), m_.clone(), v_.clone()))
    for a_, b_ in zip(*res):
        assert rel(a_, b_) < 1e-6                        # same formula; the two code paths may contract FMAs differently


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("kind", ["mse", "mae", "bayesian_tv"])
@pytest.mark.parametrize("use_lat", [False, True])
# tiled shared-memory path / streaming + band paths (W % 8 == 0) / three warps per row (lane-0 and lane-31 halo loads)
@pytest.mark.parametrize("W,tW", [(150, 152), (152, 156), (520, 524)])
def test_loss_vs_oracle(dtype, kind, use_lat, W, tW):
    _loss_case(dtype, kind, use_lat, 45, 47, W, tW)


@pytest.mark.parametrize("H,W", [(1, 8), (2, 16), (3, 8), (8, 8), (9, 24), (16, 256), (17, 264)])
def test_loss_tv_band_edges(H, W):
    """bayesian_tv band kernel at its edges: fewer rows than one 8-row band, exactly one band, one row into the next band,
    a single 8-column strip, exactly one full warp per row, one strip into the second warp."""
    for dtype in DT:
        _loss_case(dtype, "bayesian_tv", True, H, H + 1, W, W + 4)


def _loss_case(dtype, kind, use_lat, H, tH, W, tW):
    from oracle import cases, reslim_oracle as O
    from orbit2_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(3)
    B, C = 2, 3
    pred = torch.randn(B, C, H, W, generator=g).to(dtype)
    y = torch.randn(B, C, tH, tW, generator=g)
    out_vars = ["total_precipitation_24hr", "orography", "2m_temperature_max"]   # ch0 clamped, ch1 constant
    lat = np.linspace(80, -80, H)
    lw = O.lat_weights(lat) if use_lat else None
    vw = cases.VAR_WEIGHTS
    pr = pred.double().requires_grad_(True)
    yc = y.double()[:, :, :H, :W]
    yh = O.clip_replace_constant(yc, pr, out_vars)
    if kind == "mae":
        ref = O.mae(yh, yc, False, lw)
    else:
        ref = getattr(O, kind)(yh, yc, out_vars, vw, False, lw)
    ref[-1].backward()
    chw = None if kind == "mae" else torch.tensor([vw.get(v, 1.0) for v in out_vars], dtype=torch.float32, device="cuda")
    latw = torch.from_numpy(lw.reshape(-1).numpy()).float().cuda() if use_lat else None
    k = {"mse": L.LOSS_MSE, "mae": L.LOSS_MAE, "bayesian_tv": L.LOSS_BAYESIAN_TV}[kind]
    lv, dp = ops.loss_fwd_bwd(pred.cuda(), y.cuda(), k, lat_w=latw, ch_w=chw, clamp_ch=0, const_mask=0b010)
    assert rel(lv.cpu(), ref.detach()) < 1e-5
    gtol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel(dp.cpu(), pr.grad) < gtol
    # in-place clip/replace for evaluation
    pc = pred.cuda().clone()
    ops.clip_replace_(pc, y.cuda(), 0, 0b010)
    assert rel(pc.cpu(), O.clip_replace_constant(y[:, :, :H, :W], pred.float(), out_vars)) < (1e-6 if dtype == torch.float32 else 4e-3)


def test_loss_golden(golden_dir):
    """reference functional losses (float64 fixtures from the live reference)."""
    import os
    from oracle import cases
    from orbit2_b200 import _lib as L, ops
    z = np.load(os.path.join(golden_dir, "loss_vectors.npz"))
    pred = torch.from_numpy(z["pred"]).float().cuda(); tgt = torch.from_numpy(z["target"]).float().cuda()
    chw = torch.tensor([cases.VAR_WEIGHTS[v] for v in cases.OUT_VARS_3], dtype=torch.float32, device="cuda")
    from oracle import reslim_oracle as O
    latw = O.lat_weights(z["lat"]).reshape(-1).float().cuda()
    for sfx, lw in (("", None), ("_lat", latw)):
        for nm, k, cw in (("mse", L.LOSS_MSE, chw), ("bayesian_tv", L.LOSS_BAYESIAN_TV, chw), ("mae", L.LOSS_MAE, None)):
            lv, dp = ops.loss_fwd_bwd(pred, tgt, k, lat_w=lw, ch_w=cw)
            assert rel(lv.cpu(), torch.from_numpy(z[nm + sfx])) < 1e-5, nm + sfx
            assert rel(dp.cpu(), torch.from_numpy(z[nm + sfx + "_grad"])) < 1e-5, nm + sfx


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,V,gh,gw,heads,hd,extra", [(2, 8, 4, 8, 2, 64, 0), (1, 23, 9, 7, 4, 32, 1), (3, 7, 33, 5, 1, 64, 0)])
def test_frontend(dtype, B, V, gh, gw, heads, hd, extra):
    """o2_frontend_* against the same collapse written in torch (float64), incl. ignored extra rows/cols of x."""
    from orbit2_b200 import ops
    p, PP = 2, 4
    g = torch.Generator(device="cuda").manual_seed(V * gh)
    x = torch.randn(B, V, gh * p + extra, gw * p + extra, generator=g, device="cuda")
    tab_s = torch.randn(V, heads, PP + 1, generator=g, device="cuda")
    tab_v = torch.randn(heads, V * (PP + 1), hd, generator=g, device="cuda") * 0.3
    dout = torch.randn(B * gh * gw, heads * hd, generator=g, device="cuda").to(dtype)
    o = ops.frontend_fwd(x, tab_s, tab_v, p, gh, gw, hd, dtype)
    ts = tab_s.double().requires_grad_(True); tv = tab_v.double().requires_grad_(True)
    xc = x[:, :, :gh * p, :gw * p].double()
    P = xc.reshape(B, V, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, V, PP)
    P1 = torch.cat([P, torch.ones_like(P[..., :1])], -1)                     # T,V,PP+1
    sc = torch.einsum("tvk,vhk->tvh", P1, ts)
    a = sc.softmax(1)
    coef = torch.einsum("tvh,tvk->thvk", a, P1).reshape(-1, heads, V * (PP + 1))
    ref = torch.einsum("thk,hke->the", coef, tv).reshape(-1, heads * hd)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel(o, ref.detach()) < tol
    ref.backward(dout.double())
    dts, dtv = ops.frontend_bwd(x, tab_s, tab_v, dout, p, gh, gw, hd)
    gt = 2e-5 if dtype == torch.float32 else 1e-2      # bf16 arm: tensor-core contractions on bf16-rounded operands
    assert rel(dts, ts.grad) < gt
    assert rel(dtv, tv.grad) < gt


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("generic,with_g1", [(False, True), (False, False), (True, True)])
@pytest.mark.parametrize("B,C,gh,gw,extra", [(2, 3, 4, 8, 0), (1, 3, 6, 12, 1), (1, 1, 5, 21, 0), (2, 3, 12, 20, 0),
                                             (1, 2, 9, 17, 2)])
def test_conv1_headtail(dtype, B, C, gh, gw, extra, generic, with_g1, monkeypatch):
    """residual branch + head tail against the oracle's path2 / unpatchify / conv_out (float64); both the register-tiled
    kernels (p 2 / mag 4 / cr 4: every reference config) and the generic ones (forced by O2_HEADTAIL_GENERIC=1)."""
    from oracle import reslim_oracle as O
    from orbit2_b200 import ops
    monkeypatch.setenv("O2_HEADTAIL_GENERIC", "1" if generic else "0")
    p, mag, cr = 2, 4, 4
    V = C + 6
    Hx, Wx = gh * p + extra, gw * p + extra
    g = torch.Generator(device="cuda").manual_seed(gh * gw)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
    x = rn(B, V, Hx, Wx)
    idx = list(range(V - 1, V - 1 - (C + 4), -1))
    w1, b1 = rn(cr * mag * mag, C + 4, 3, 3) * 0.2, rn(cr * mag * mag) * 0.1
    w2, b2 = rn(C, cr, 3, 3) * 0.3, rn(C) * 0.1
    wo, bo = rn(C, C, 3, 3) * 0.3, rn(C) * 0.1
    ho = rn(B * gh * gw, C * (mag * p) ** 2).to(dtype)
    Ho, Wo = gh * p * mag, gw * p * mag
    dp = rn(B, C, Ho, Wo).to(dtype)
    if with_g1:
        h1, g1 = ops.path2_conv1_fwd(x, idx, w1, b1, dtype, mag=mag)
        gref = torch.nn.functional.pixel_shuffle(torch.nn.functional.gelu(h1.double()), mag)
        assert rel(g1, gref) < (1e-6 if dtype == torch.float32 else 5e-3)
    else:
        h1, g1 = ops.path2_conv1_fwd(x, idx, w1, b1, dtype), None
    preds = ops.headtail_fwd(ho, h1, wo, bo, w2, b2, B, C, gh, gw, p, mag, g1=g1)
    d = lambda t: t.double().requires_grad_(True)
    w1d, b1d, w2d, b2d, wod, bod, hod = d(w1), d(b1), d(w2), d(b2), d(wo), d(bo), d(ho)
    sd = {"path2.0.weight": w1d, "path2.0.bias": b1d, "path2.3.weight": w2d, "path2.3.bias": b2d}
    p2 = O.path2(sd, x[:, idx].double(), mag)
    img = O.unpatchify(hod.reshape(B, gh * gw, -1), (gh * p, gw * p), p, mag, C)
    img = torch.nn.functional.conv2d(img, wod, bod, padding=1)
    ref = img + p2[:, :, :Ho, :Wo]
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel(preds, ref.detach()) < tol
    ref.backward(dp.double())
    G = {k: torch.zeros_like(v) for k, v in dict(w1=w1, b1=b1, w2=w2, b2=b2, wo=wo, bo=bo).items()}
    dho, dh1 = ops.headtail_bwd(dp, ho, h1, wo, w2, G["wo"], G["bo"], G["w2"], G["b2"], B, C, gh, gw, p, mag, g1=g1)
    ops.path2_conv1_bwd(x, idx, dh1, G["w1"], G["b1"])
    gt = 2e-5 if dtype == torch.float32 else 2.5e-2
    assert rel(dho, hod.grad) < gt
    for k, r in dict(wo=wod, bo=bod, w2=w2d, b2=b2d, w1=w1d, b1=b1d).items():
        assert rel(G[k], r.grad) < gt, k


def _attn_ref(qkv, B, N, heads, hd):
    t = qkv.double().reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4).requires_grad_(True)
    q, k, v = t.unbind(0)
    s = (q * hd ** -0.5) @ k.transpose(-2, -1)
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * N, heads * hd)
    return t, o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,N,heads,scale_in", [(1, 100, 1, 0.5), (2, 512, 2, 0.5), (1, 333, 1, 0.5), (1, 64, 2, 0.5),
                                                (2, 129, 1, 0.5), (1, 1000, 2, 1.5), (3, 40, 3, 0.5)])
def test_attn_hd256(dtype, B, N, heads, scale_in):
    """head dim 256 (interm_10b: 32 heads x 256), forward and backward vs float64 attention on the same operands.
    bf16: the tcgen05 kernels of csrc/attn_tc256.cu (owner tile of 128 rows, 64-row streamed steps: N covers < one step,
    exactly one step, one past an owner tile, ragged tails, 512 = the 10b grid; scale_in 1.5 gives logits large enough
    that the running maximum jumps by more than 2^8 between steps -> the lazy O-rescale path).  fp32: the chunked SIMT
    arm (also what a bf16 call with attention dropout runs on)."""
    from orbit2_b200 import ops
    hd = 256
    g = torch.Generator(device="cuda").manual_seed(N + 11 * heads)
    D = heads * hd
    qkv = (torch.randn(B * N, 3 * D, generator=g, device="cuda") * scale_in).to(dtype)
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(dtype)
    out, lse = ops.attn_fwd(qkv, B, N, heads, hd)
    t, o, lse_ref = _attn_ref(qkv, B, N, heads, hd)
    assert out.dtype == dtype and rel(lse, lse_ref.detach()) < (1e-4 if dtype == torch.float32 else 1e-3)
    assert rel(out, o.detach()) < (2e-5 if dtype == torch.float32 else 1.5e-2)
    o.backward(dout.double())
    dqkv = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd)
    ref = t.grad.permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D)
    gtol = 5e-5 if dtype == torch.float32 else 2e-2
    assert dqkv.dtype == dtype and rel(dqkv, ref) < gtol
    for i, name in enumerate(("dq", "dk", "dv")):                   # each gradient against its own scale
        assert rel(dqkv.view(B * N, 3, D)[:, i], ref.view(B * N, 3, D)[:, i]) < gtol, name


def test_attn_hd256_dropout_runs_on_fp32_arm():
    """bf16 + attention dropout at head dim 256 has no tcgen05 variant: ops routes it through the fp32 arm and the C-ABI
    tcgen05 entry refuses it loudly instead of silently ignoring the mask."""
    from oracle import dropout_mask as DM
    from orbit2_b200 import _lib, ops
    B, N, heads, hd, p = 1, 150, 1, 256, 0.25
    seed, site = 99, 3
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = (torch.randn(B * N, 3 * hd, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, B, N, heads, hd, (p, seed, site))
    M = DM.attn_scaled_mask(seed, site, B, heads, N, p).cuda()
    t = qkv.double().reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = t.unbind(0)
    o = (((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(-1) * M @ v).transpose(1, 2).reshape(B * N, hd)
    assert rel(out, o) < 1.5e-2
    lib = _lib.load()
    rc = lib.o2_attn_fwd_drop(ops.GEMM_TC_BF16, ops._ptr(qkv), ops._ptr(out), ops._ptr(lse), B, N, heads, hd, hd ** -0.5,
                              p, seed, site, ops._stream())
    assert rc != 0


@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("B,N,heads,scale_in", [(1, 128, 1, 1.0), (2, 256, 2, 1.0), (1, 300, 2, 2.0), (2, 1000, 3, 1.0),
                                                (1, 72, 1, 4.0), (1, 2049, 1, 1.0), (1, 40, 2, 1.0), (2, 1500, 1, 1.0)])
def test_attn_tc(B, N, heads, scale_in, two_pass):
    """tcgen05 flash attention (bf16) vs float64 softmax attention on the same bf16-rounded operands; N covers the
    ragged tails (N % 128 = 72 like the 117M grid's 16200, < one tile, < one sub-tile, one past a tile, an odd number of
    64-query sub-tiles) and large logits.  Backward: the one-pass kernel (dQ partials reduced by TMA into an fp32
    accumulator) and the deterministic two-kernel path."""
    from orbit2_b200 import ops
    hd = 64
    g = torch.Generator(device="cuda").manual_seed(N + heads)
    D = heads * hd
    qkv = (torch.randn(B * N, 3 * D, generator=g, device="cuda") * scale_in).to(torch.bfloat16)
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, B, N, heads, hd)
    t, o, lse_ref = _attn_ref(qkv, B, N, heads, hd)
    assert rel(lse, lse_ref.detach()) < 1e-3
    assert rel(out, o.detach()) < 1.5e-2
    o.backward(dout.double())
    dqkv = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd, two_pass=two_pass)
    ref = t.grad.permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D)
    assert rel(dqkv, ref) < 2e-2
    for i, name in enumerate(("dq", "dk", "dv")):                   # each gradient against its own scale
        assert rel(dqkv.view(B * N, 3, D)[:, i], ref.view(B * N, 3, D)[:, i]) < 2e-2, name


def test_attn_bwd_one_pass_vs_two_pass():
    """The one-pass backward against the deterministic two-kernel backward on the same inputs: dK / dV are single-owner
    tensor-memory sums in both (same bf16 P^T / dS^T operands, same order) and must agree to bf16 rounding of the final
    store; dQ is a sum of per-key-tile fp32 partials reduced in the L2 (order not fixed) instead of one TMEM accumulation:
    equal to fp32 rounding, i.e. far inside one bf16 ulp of the result for almost every element.  The two-kernel path is
    bit-reproducible run to run."""
    from orbit2_b200 import ops
    B, N, heads, hd = 2, 1100, 3, 64
    g = torch.Generator(device="cuda").manual_seed(77)
    D = heads * hd
    qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda").to(torch.bfloat16)
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, B, N, heads, hd)
    a = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd, two_pass=True)
    a2 = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd, two_pass=True)
    assert torch.equal(a, a2)
    f = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd, two_pass=False)
    av, fv = a.float().view(B * N, 3, D), f.float().view(B * N, 3, D)
    for i, name in enumerate(("dq", "dk", "dv")):
        assert rel(fv[:, i], av[:, i]) < 8e-3, name                  # <= one bf16 ulp at the top of the range
    assert ((fv[:, 0] - av[:, 0]).abs() > 0).float().mean().item() < 0.5


@pytest.mark.parametrize("B,N,heads", [(1, 128, 1), (2, 300, 2), (1, 1000, 3), (1, 72, 1), (1, 2049, 2)])
def test_attn_tc_hd128(B, N, heads):
    """head dim 128 (interm_1b: 24 heads x 128): tcgen05 forward and backward vs float64 on the bf16-rounded operands."""
    from orbit2_b200 import ops
    hd = 128
    g = torch.Generator(device="cuda").manual_seed(N * 7 + heads)
    D = heads * hd
    qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda").to(torch.bfloat16)
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(torch.bfloat16)
    out, lse = ops.attn_fwd(qkv, B, N, heads, hd)
    t, o, lse_ref = _attn_ref(qkv, B, N, heads, hd)
    assert rel(lse, lse_ref.detach()) < 1e-3
    assert rel(out, o.detach()) < 1.5e-2
    dqkv = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd)
    o.backward(dout.double())
    ref = t.grad.permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D)
    assert rel(dqkv, ref) < 2e-2


@pytest.mark.parametrize("dtype,hd,B,N,heads", [(torch.bfloat16, 64, 2, 300, 2), (torch.bfloat16, 64, 1, 1000, 1),
                                                (torch.bfloat16, 128, 1, 333, 2), (torch.float32, 64, 2, 200, 2),
                                                (torch.float32, 32, 1, 77, 3)])
@pytest.mark.parametrize("p", [0.1, 0.35])
@pytest.mark.parametrize("two_pass", [False, True])
def test_attn_dropout(dtype, hd, B, N, heads, p, two_pass):
    """Attention-probability dropout inside the attention kernels (forward, dQ, dK/dV; tcgen05 and fp32 SIMT arms) against
    float64 attention with the SAME mask rebuilt from the documented hash (oracle/dropout_mask.py)."""
    from oracle import dropout_mask as DM
    from orbit2_b200 import ops
    seed, site = 0x1234_5678_9ABC_DEF0, 17
    g = torch.Generator(device="cuda").manual_seed(N + hd)
    D = heads * hd
    qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda").to(dtype)
    dout = torch.randn(B * N, D, generator=g, device="cuda").to(dtype)
    out, lse = ops.attn_fwd(qkv, B, N, heads, hd, (p, seed, site))
    M = DM.attn_scaled_mask(seed, site, B, heads, N, p).cuda()
    if two_pass and not (dtype == torch.bfloat16 and hd == 64):
        pytest.skip("only the bf16 head-dim-64 arm has two backward paths")
    assert abs(float((M > 0).double().mean()) - (1 - p)) < 0.02
    t = qkv.double().reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4).requires_grad_(True)
    q, k, v = t.unbind(0)
    s = (q * hd ** -0.5) @ k.transpose(-2, -1)
    o = ((s.softmax(-1) * M) @ v).transpose(1, 2).reshape(B * N, D)
    o.backward(dout.double())
    ref = t.grad.permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D)
    ftol, gtol = (1.5e-2, 2e-2) if dtype == torch.bfloat16 else (2e-5, 5e-5)
    assert rel(lse, torch.logsumexp(s, -1).detach()) < 1e-3
    assert rel(out, o.detach()) < ftol
    dqkv = ops.attn_bwd(qkv, out, dout, lse, B, N, heads, hd, (p, seed, site), two_pass=two_pass)
    assert rel(dqkv, ref) < gtol
    # p = 0 through the same entry point is the plain kernel
    out0, _ = ops.attn_fwd(qkv, B, N, heads, hd, (0.0, seed, site))
    out1, _ = ops.attn_fwd(qkv, B, N, heads, hd)
    assert torch.equal(out0, out1)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("p", [0.1, 0.35])
def test_attn_dropout_mask_bits(dtype, p):
    """The keep mask of the attention kernels read back bit by bit: with Q = K = 0 the probabilities are uniform (1 / N)
    and with V = identity the output row q IS the masked probability row, so out * N > 0.5 <=> keep(q, k).  Compared with
    the restatement in oracle/dropout_mask.py (itself pinned by tests/golden/keep_masks.npz) for both arms."""
    from oracle import dropout_mask as DM
    from orbit2_b200 import ops
    seed, site = 0x1234_5678_9ABC_DEF0, 17
    B, N, heads, hd = 2, 64, 3, 64
    qkv = torch.zeros(B, N, 3, heads, hd, device="cuda")
    qkv[:, torch.arange(N), 2, :, torch.arange(N)] = 1.0              # V[b, n, h, :] = e_n
    qkv = qkv.reshape(B * N, 3 * heads * hd).to(dtype)
    out, _ = ops.attn_fwd(qkv, B, N, heads, hd, (p, seed, site))
    got = out.float().reshape(B, N, heads, hd).permute(0, 2, 1, 3)[..., :N] * N
    M = DM.attn_scaled_mask(seed, site, B, heads, N, p).cuda()
    assert torch.equal(got > 0.5, M > 0)
    kept = got[got > 0.5]
    assert (kept - 65536.0 / (65536 - int(p * 65536))).abs().max().item() < (1e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("ih,iw,oh,ow,D", [(4, 8, 6, 12, 64), (8, 16, 5, 11, 128), (45, 90, 49, 94, 256), (3, 6, 24, 48, 32),
                                           (16, 32, 16, 30, 1024), (2, 4, 1, 2, 8)])
def test_bicubic_resample(ih, iw, oh, ow, D):
    """o2_bicubic_fwd / _bwd (channels-last pos_embed resample) vs torch's upsample_bicubic2d and its autograd adjoint on
    the same fp32 table: up- and down-sampling, non-integer ratios, magnification 8 (many outputs per clamped border tap)."""
    from orbit2_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(ih * 131 + ow)
    src = torch.randn(ih * iw, D, generator=g, device="cuda")
    d = torch.randn(oh * ow, D, generator=g, device="cuda")
    t = src.double().reshape(1, ih, iw, D).permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.interpolate(t, size=(oh, ow), mode="bicubic", align_corners=False)
    ref.backward(d.double().reshape(1, oh, ow, D).permute(0, 3, 1, 2))
    # float64 torch: the tap weights differ by the fp32 rounding of the source coordinate (both torch's fp32 kernel and
    # ours compute scale * (dst + 0.5) - 0.5 in fp32, as the reference does); fp32 torch on the same device: rounding order only
    out = ops.bicubic_fwd(src, ih, iw, oh, ow)
    assert rel(out, ref.permute(0, 2, 3, 1).reshape(oh * ow, D)) < 3e-5
    t32 = src.reshape(1, ih, iw, D).permute(0, 3, 1, 2).clone().requires_grad_(True)
    ref32 = F.interpolate(t32, size=(oh, ow), mode="bicubic", align_corners=False)
    ref32.backward(d.reshape(1, oh, ow, D).permute(0, 3, 1, 2))
    assert rel(out, ref32.permute(0, 2, 3, 1).reshape(oh * ow, D)) < 3e-6
    dsrc = ops.bicubic_bwd(d, ih, iw, oh, ow)
    assert rel(dsrc, t.grad.permute(0, 2, 3, 1).reshape(ih * iw, D)) < 3e-5
    assert rel(dsrc, t32.grad.permute(0, 2, 3, 1).reshape(ih * iw, D)) < 1e-5      # torch's backward scatters with atomics
    assert torch.equal(dsrc, ops.bicubic_bwd(d, ih, iw, oh, ow))          # gather form: bit-reproducible


def test_pos_embed_resample_matches_oracle():
    """Res_Slim_ViT.pos_res_embed on a grid that differs from the stored one (pos_embed.py:103-138) vs the oracle's
    F.interpolate restatement, value and pos_embed gradient."""
    from oracle import cases, reslim_oracle as O
    from tests.util import build_model
    cfg = cases.get_case("tiny")
    m = build_model(cfg).cuda()
    m.img_size = (12, 24)                              # grid differs from the 8x16 init grid -> bicubic resample
    with torch.no_grad():
        m.pos_embed.normal_()
        m.spatial_embed.bias.zero_()
    m.spatial_resolution = 0.0
    pe = m.pos_embed.detach().cpu().double().requires_grad_(True)
    ref = O.interp_pos_embed(pe, 2, (12, 24))[0]
    out = m.pos_res_embed(6, 12, torch.float32)
    assert rel(out.cpu(), ref) < 3e-5                  # float64 oracle vs fp32 source coordinates
    ref32 = O.interp_pos_embed(m.pos_embed.detach().cpu(), 2, (12, 24))[0]
    assert rel(out.cpu(), ref32) < 3e-6
    w = torch.randn(ref.shape, dtype=torch.float64)
    (ref * w).sum().backward()
    (out * w.float().cuda()).sum().backward()
    assert rel(m.pos_embed.grad.cpu(), pe.grad) < 3e-5


def test_dropout_step_word_masks():
    """o2_dropout_seed_source: with a device-resident step word installed, the token-stream and attention masks are the
    restatement's masks for (seed, site, word) -- the word is read by the KERNEL, so rewriting it changes the next launch
    with identical host arguments (what a CUDA-graph replay does) -- and removing it restores the by-value masks."""
    from oracle import dropout_mask as DM
    from orbit2_b200 import ops
    seed, site, p = 0x0FED_CBA9_8765_4321, 21, 0.3
    y = torch.ones(40, 64, device="cuda")
    word = torch.zeros(1, device="cuda", dtype=torch.int64)
    B, N, heads, hd = 1, 64, 2, 64
    qkv = torch.zeros(B, N, 3, heads, hd, device="cuda")
    qkv[:, torch.arange(N), 2, :, torch.arange(N)] = 1.0
    qkv = qkv.reshape(B * N, 3 * heads * hd).to(torch.bfloat16)
    plain = ops.dropout(y, p, seed, site)
    try:
        ops.dropout_seed_source(word)
        seen = []
        for w in (0, 0x1111_2222_3333_4444, -5):
            word.fill_(w)
            got = ops.dropout(y, p, seed, site)
            ref = DM.keep_mask(seed, site, y.numel(), p, step_word=w).view_as(y).cuda()
            assert torch.equal(got > 0, ref), hex(w & 0xFFFFFFFFFFFFFFFF)
            out, _ = ops.attn_fwd(qkv, B, N, heads, hd, (p, seed, site))
            am = out.float().reshape(B, N, heads, hd).permute(0, 2, 1, 3)[..., :N] * N > 0.5
            assert torch.equal(am, DM.attn_scaled_mask(seed, site, B, heads, N, p, step_word=w).cuda() > 0)
            # the GEMM epilogue reads the same word
            a = torch.ones(40, 64, device="cuda", dtype=torch.bfloat16)
            wgt = torch.eye(64, device="cuda", dtype=torch.bfloat16)
            c = ops.gemm(a, wgt, torch.empty(40, 64, device="cuda", dtype=torch.bfloat16), epi=ops.EPI_BIAS_RES,
                         bias=torch.zeros(64, device="cuda"), aux=torch.zeros(40, 64, device="cuda", dtype=torch.bfloat16),
                         drop=(p, seed, site, None, 0))
            assert torch.equal(c > 0, ref)
            seen.append(got > 0)
        assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])
    finally:
        ops.dropout_seed_source(None)
    assert torch.equal(ops.dropout(y, p, seed, site), plain)
    assert torch.equal(plain > 0, DM.keep_mask(seed, site, y.numel(), p).view_as(y).cuda())


@pytest.mark.parametrize("T,D,rps", [(1000, 1024, 250), (129, 128, 129), (67, 512, 10), (300, 3072, 100)])
def test_layernorm_bwd_masked_second_output(T, D, rps):
    """o2_layernorm_bwd_drop: the second output is bit for bit what o2_dropout makes of dx (same hash mask, drop-path factor
    per sample); dx, dgamma, dbeta are those of the plain call.  D = 3072 takes the separate-pass fallback of ops.layernorm_bwd."""
    from orbit2_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(T + D)
    bf = torch.bfloat16
    x = (torch.randn(T, D, generator=g, device="cuda") * 2 + 0.5).to(bf)
    gamma = torch.randn(D, generator=g, device="cuda")
    beta = torch.randn(D, generator=g, device="cuda")
    dy = torch.randn(T, D, generator=g, device="cuda").to(bf)
    dres = torch.randn(T, D, generator=g, device="cuda").to(bf)
    ss = (torch.rand((T + rps - 1) // rps, generator=g, device="cuda") < 0.8).float() / 0.8
    _, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    dg0, db0 = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    dx0 = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dg0, db0, dres=dres)
    for p, scale in ((0.2, ss), (0.0, ss), (0.3, None)):
        dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
        dx, dxm = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dg, db, dres=dres, drop=(p, 0xABCDEF0123, 9, scale, rps))
        assert torch.equal(dx, dx0)
        assert torch.equal(dxm, ops.dropout(dx0, p, 0xABCDEF0123, 9, sample_scale=scale, rows_per_sample=rps if scale is not None else 0))
        assert rel(dg, dg0) < 1e-5 and rel(db, db0) < 1e-5            # column sums: atomics order only
