#!/usr/bin/env python
"""The "kernel to beat" of SURVEY.md headline 1 / BASELINE.md section 4: the reference's schedule, as written, on the SAME
B200 -- PyTorch eager + F.scaled_dot_product_attention (components/attention.py:66-71, FusedAttn.DEFAULT) + cuBLASLt.

The Python reference cannot travel to the GPU box, so this tool runs its functional restatement (oracle/reslim_oracle.py,
the per-variable patch embeds, the [B,V,L,D] stack and the 1.56 TFLOP variable-aggregation kv GEMM included) with
``.cuda()`` tensors in bf16 (autocast-free: parameters cast once, like FSDP MixedPrecision bf16) and in fp32 (TF32 off and
on), forward + clip + bayesian_tv + backward, and prints one JSON line per configuration.  It also times the library
attention alone (SDPA flash / cuDNN, forward and backward at L = 16200, 16 heads x 64) next to this repo's kernels.
TEST / MEASUREMENT TOOL: nothing in the product imports it.

    python tools/ref_gpu_bench.py [--batch 8] [--steps 3] > profiles/r02_ref_gpu.jsonl
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def event_ms(fn, iters, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def whole_step(case, B, dtype, tf32, steps):
    from oracle import cases, reslim_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    O.USE_SDPA = True
    cfg = cases.get_case(case)
    sd = {k: v.cuda().to(dtype).requires_grad_(True) for k, v in O.init_state_dict(cfg, 0).items()}
    x, y = O.synthetic_batch(cfg, B, cfg["in_vars"], cfg["out_vars"], 0)
    x, y = x.cuda().to(dtype), y.cuda()

    def step():
        loss = O.training_step(sd, cfg, x, y, cfg["in_vars"], cfg["out_vars"], "bayesian_tv", cfg["var_weights"])
        loss.backward()
        for v in sd.values():
            v.grad = None
        return loss

    torch.cuda.reset_peak_memory_stats()
    ms = event_ms(step, steps, warm=1)
    return {"what": "reference schedule (oracle port, eager + SDPA + cuBLASLt) fwd+clip+bayesian_tv+bwd on this GPU",
            "workload": case, "B": B, "dtype": str(dtype).replace("torch.", ""), "tf32": tf32, "ms_per_step": ms,
            "samples_per_s": B / (ms * 1e-3), "hbm_peak_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}


def attention_only(B, N, heads, hd, iters, p_drop=0.0):
    """Library SDPA (the back ends torch picks for bf16: flash / cuDNN / mem-efficient) vs this repo's tcgen05 kernels."""
    from torch.nn.attention import SDPBackend, sdpa_kernel
    from orbit2_b200 import ops
    out = []
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * N, 3 * heads * hd, generator=g, device="cuda").to(torch.bfloat16)
    dout = torch.randn(B * N, heads * hd, generator=g, device="cuda").to(torch.bfloat16)
    fl_fwd = 4.0 * N * N * heads * hd * B
    q, k, v = (t.contiguous().requires_grad_(True) for t in qkv.view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4).unbind(0))
    do4 = dout.view(B, N, heads, hd).transpose(1, 2).contiguous()
    for name, be in (("flash", SDPBackend.FLASH_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION),
                     ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel([be]):
                f_ms = event_ms(lambda: F.scaled_dot_product_attention(q, k, v, dropout_p=p_drop), iters)
                o = F.scaled_dot_product_attention(q, k, v, dropout_p=p_drop)

                def bwd():
                    for t in (q, k, v):
                        t.grad = None
                    o.backward(do4, retain_graph=True)
                b_ms = event_ms(bwd, iters)
            out.append({"what": f"torch SDPA backend {name}", "dropout_p": p_drop, "B": B, "N": N, "heads": heads, "hd": hd, "fwd_ms": f_ms,
                        "bwd_ms": b_ms, "fwd_tflops": fl_fwd / f_ms / 1e9, "bwd_tflops_algorithmic": 2.5 * fl_fwd / b_ms / 1e9})
        except Exception as e:                                    # a back end that is not built / not eligible here
            out.append({"what": f"torch SDPA backend {name}", "unavailable": repr(e)[:200]})
    dr = (p_drop, 1234567, 3) if p_drop > 0 else None
    o2, lse = ops.attn_fwd(qkv, B, N, heads, hd, dr)
    f_ms = event_ms(lambda: ops.attn_fwd(qkv, B, N, heads, hd, dr), iters)
    b_ms = event_ms(lambda: ops.attn_bwd(qkv, o2, dout, lse, B, N, heads, hd, dr), iters)
    out.append({"what": "orbit2_b200 tcgen05 attention (this repo)", "dropout_p": p_drop, "B": B, "N": N, "heads": heads, "hd": hd, "fwd_ms": f_ms,
                "bwd_ms": b_ms, "fwd_tflops": fl_fwd / f_ms / 1e9, "bwd_tflops_algorithmic": 2.5 * fl_fwd / b_ms / 1e9})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--workload", default="117m")
    ap.add_argument("--skip-step", action="store_true")
    ap.add_argument("--drop", type=float, default=0.0, help="attention-probability dropout for the attention-only comparison")
    a = ap.parse_args()
    for row in attention_only(a.batch, 16200, 16, 64, 3, a.drop):
        print(json.dumps(row), flush=True)
    if a.skip_step:
        return
    for dtype, tf32 in ((torch.bfloat16, False), (torch.float32, True), (torch.float32, False)):
        B = a.batch
        while B >= 1:
            try:
                print(json.dumps(whole_step(a.workload, B, dtype, tf32, a.steps)), flush=True)
                break
            except torch.OutOfMemoryError:
                torch.cuda.empty_cache()
                print(json.dumps({"what": "reference schedule", "B": B, "dtype": str(dtype), "oom": True}), flush=True)
                B //= 2


if __name__ == "__main__":
    main()
