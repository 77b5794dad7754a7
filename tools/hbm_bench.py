"""Achieved HBM bandwidth of the memory-bound kernels at the interm_117m shapes (B=8): algorithmic bytes (DESIGN.md
section 3) / CUDA-event time, against the measured copy peak in MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from orbit2_b200 import _lib as L, ops  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    SRC = "measured"
except Exception:
    PEAK, SRC = 6650.0, "fallback"
B, V, C, H, W, p, mag, heads, hd, D = 8, 23, 3, 180, 360, 2, 4, 16, 64, 1024
gh, gw = H // p, W // p
T = B * gh * gw
Ho, Wo = H * mag, W * mag
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device=dev)


def timed(fn, n=10, graph=True):
    """Mean device time of fn: n back-to-back calls captured in one CUDA graph (so a 40 us kernel is not paced by the
    Python / ctypes issue rate), replayed three times; eager event timing if the capture fails."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph and os.environ.get("O2_HBM_EAGER", "0") != "1":
        try:
            st = torch.cuda.Stream()
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                fn()
            torch.cuda.current_stream().wait_stream(st)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(n):
                    fn()
            gr.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / (3 * n)
        except Exception as ex:          # noqa: BLE001
            print(f"# graph capture failed ({type(ex).__name__}: {ex}); eager timing", file=sys.stderr)
            torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []


def report(name, fn, nbytes, graph=True):
    ms = timed(fn, graph=graph)
    gbs = nbytes / ms / 1e6
    rows.append((name, ms, nbytes / 1e6, gbs, gbs / PEAK))
    print(f"{name:34s} {ms:8.3f} ms {nbytes / 1e6:9.1f} MB {gbs:8.0f} GB/s {100 * gbs / PEAK:5.1f}% of {SRC} peak {PEAK:.0f}")


x = rn(B, V, H, W)
xt, dy, dres = rn(T, D).to(bf), rn(T, D).to(bf), rn(T, D).to(bf)
gam, bet = rn(D), rn(D)
y, mean, rstd = ops.layernorm_fwd(xt, gam, bet)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
report("layernorm fwd [T,1024] bf16", lambda: ops.layernorm_fwd(xt, gam, bet), 2 * T * D * 2)
report("layernorm bwd (+residual grad)", lambda: ops.layernorm_bwd(dy, xt, gam, mean, rstd, dg, db, dres=dres), 4 * T * D * 2)
pred, tgt = rn(B, C, Ho, Wo).to(bf), rn(B, C, Ho, Wo)
chw = torch.tensor([1.0, 10.0, 10.0], device=dev)
report("loss bayesian_tv fwd+grad", lambda: ops.loss_fwd_bwd(pred, tgt, L.LOSS_BAYESIAN_TV, ch_w=chw, clamp_ch=0),
       B * C * Ho * Wo * (2 + 4 + 2))
report("loss mse fwd+grad", lambda: ops.loss_fwd_bwd(pred, tgt, L.LOSS_MSE, ch_w=chw, clamp_ch=0), B * C * Ho * Wo * (2 + 4 + 2))
idx = [21, 6, 5, 0, 1, 2, 3]
w1, b1 = rn(64, 7, 3, 3) * 0.1, rn(64) * 0.1
w2, b2, wo, bo = rn(C, 4, 3, 3) * 0.1, rn(C) * 0.1, rn(C, C, 3, 3) * 0.1, rn(C) * 0.1
h1, g1 = ops.path2_conv1_fwd(x, idx, w1, b1, bf, mag=mag)
ho = rn(T, C * (mag * p) ** 2).to(bf)
report("path2 conv1 fwd (7->64 ch, low-res)", lambda: ops.path2_conv1_fwd(x, idx, w1, b1, bf, mag=mag), B * 7 * H * W * 4 + 2 * B * 64 * H * W * 2)
report("head tail fwd (unpatchify+convs+add)", lambda: ops.headtail_fwd(ho, h1, wo, bo, w2, b2, B, C, gh, gw, p, mag, g1=g1),
       ho.numel() * 2 + h1.numel() * 2 + B * C * Ho * Wo * 2)
dp = rn(B, C, Ho, Wo).to(bf)
G = [torch.zeros_like(t) for t in (wo, bo, w2, b2)]
report("head tail bwd", lambda: ops.headtail_bwd(dp, ho, h1, wo, w2, *G, B, C, gh, gw, p, mag, g1=g1),
       dp.numel() * 2 + 2 * ho.numel() * 2 + 3 * h1.numel() * 2)
dh1 = torch.randn_like(h1)
Gw = [torch.zeros_like(w1), torch.zeros_like(b1)]
report("path2 conv1 bwd (weight grads)", lambda: ops.path2_conv1_bwd(x, idx, dh1, *Gw), B * 7 * H * W * 4 + B * 64 * H * W * 2)
ts, tv = rn(V, heads, 5), rn(heads, V * 5, hd) * 0.1
do = rn(T, D).to(bf)
report("front end fwd (x -> o [T,D])", lambda: ops.frontend_fwd(x, ts, tv, p, gh, gw, hd, bf), x.numel() * 4 + T * D * 2)
report("front end bwd (table grads)", lambda: ops.frontend_bwd(x, ts, tv, do, p, gh, gw, hd), x.numel() * 4 + T * D * 2)
cs = torch.zeros(D, device=dev)
report("colsum [T,1024] bf16 (bias grad)", lambda: ops.colsum(xt, cs), T * D * 2)
n = 126_100_000
P, Gd, M, Vv = (torch.zeros(n, device=dev) for _ in range(4))
Pb = torch.zeros(n, device=dev, dtype=bf)
report("fused AdamW 126.1M params", lambda: ops.adamw(P, Gd, M, Vv, Pb, 1e-3, 0.9, 0.99, 1e-8, 1e-5, 1), n * (4 * 4 + 3 * 4 + 2))
a, b_ = torch.empty(1 << 29, device=dev, dtype=bf), torch.empty(1 << 29, device=dev, dtype=bf)
report("torch copy 1 GiB bf16 (yardstick)", lambda: b_.copy_(a), 2 * a.numel() * 2, graph=False)
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        f.write(f"| kernel | ms | algorithmic MB | GB/s | of {SRC} HBM peak ({PEAK:.0f} GB/s) |\n|---|---|---|---|---|\n")
        for nm, ms, mb, gbs, fr in rows:
            f.write(f"| {nm} | {ms:.3f} | {mb:.1f} | {gbs:.0f} | {100 * fr:.1f}% |\n")
