"""GPU parity: tcgen05 bf16 GEMM and SIMT fp32 GEMM vs torch fp32 matmul of the same (bf16-rounded) inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return (a.double() - b.double()).abs().max().item() / (b.double().abs().max().item() + 1e-12)


def _mk(shape, dtype, g):
    return (torch.randn(shape, generator=g, device="cuda") * 0.5).to(dtype)


def _ref(a, b, ta, tb):
    A = a.float().t() if ta else a.float()
    Bm = b.float() if tb else b.float().t()
    return A @ Bm


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (200, 192, 320), (1000, 1024, 1024), (384, 128, 192)])
def test_gemm_plain(dtype, ta, tb, M, N, K):
    from orbit2_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _mk((K, M) if ta else (M, K), dtype, g)
    b = _mk((K, N) if tb else (N, K), dtype, g)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=dtype)
    ops.gemm(a, b, out, trans_a=ta, trans_b=tb)
    torch.cuda.synchronize()
    ref = _ref(a, b, ta, tb)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    err = (out.float() - ref).abs().max().item() / (ref.abs().max().item() + 1e-9)
    assert err < tol, f"rel err {err}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_epilogues(dtype):
    from orbit2_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 300, 512, 256
    a = _mk((M, K), dtype, g); w = _mk((N, K), dtype, g)
    bias = torch.randn(N, generator=g, device="cuda")
    acc = a.float() @ w.float().t()
    tol = 1e-5 if dtype == torch.float32 else 1.5e-2

    def close(x, y):
        return (x.float() - y).abs().max().item() / (y.abs().max().item() + 1e-9) < tol

    out = torch.empty(M, N, device="cuda", dtype=dtype)
    ops.gemm(a, w, out, epi=ops.EPI_BIAS, bias=bias)
    assert close(out, acc + bias)
    u = torch.empty(M, N, device="cuda", dtype=dtype)
    ops.gemm(a, w, out, epi=ops.EPI_BIAS_GELU, bias=bias, aux_out=u)
    assert close(u, acc + bias)
    assert close(out, torch.nn.functional.gelu(acc + bias))
    res = _mk((M, N), dtype, g)
    ops.gemm(a, w, out, epi=ops.EPI_BIAS_RES, bias=bias, aux=res)
    assert close(out, acc + bias + res.float())
    pos = _mk((100, N), dtype, g)       # broadcast rows (pos-embed style): row m uses pos[m % 100]
    ops.gemm(a, w, out, epi=ops.EPI_BIAS_RES, bias=bias, aux=pos, aux_rows=100)
    idx = torch.arange(M, device="cuda") % 100
    assert close(out, acc + bias + pos.float()[idx])
    pre = _mk((M, N), dtype, g)
    ops.gemm(a, w, out, epi=ops.EPI_DGELU, aux=pre)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    assert close(out, acc * x.grad)
    # fp32 output from bf16 operands, and split-K accumulation into a pre-filled buffer
    c32 = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(a, w, c32)
    assert (c32 - acc).abs().max().item() / acc.abs().max().item() < (1e-5 if dtype == torch.float32 else 2e-3)
    base = torch.randn(M, N, generator=g, device="cuda")
    c32 = base.clone()
    ops.gemm(a, w, c32, epi=ops.EPI_ACCUM, split_k=3)
    assert (c32 - (acc + base)).abs().max().item() / acc.abs().max().item() < (1e-5 if dtype == torch.float32 else 2e-3)


def test_gemm_wgrad_shape():
    """dW[N,K] = dY^T X with a long contraction (tokens) and split-K, bf16."""
    from orbit2_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    T, N, K = 4096 + 72, 384, 256
    dy = _mk((T, N), torch.bfloat16, g); x = _mk((T, K), torch.bfloat16, g)
    dw = torch.zeros(N, K, device="cuda")
    ops.gemm(dy, x, dw, trans_a=True, trans_b=True, epi=ops.EPI_ACCUM, split_k=8)
    ref = dy.float().t() @ x.float()
    assert (dw - ref).abs().max().item() / ref.abs().max().item() < 2e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T,N,K,split", [(4096 + 72, 384, 256, 8), (1000, 128, 512, 1), (5000, 200, 1024, 3),
                                         (129600 // 8, 1024, 1024, 5), (333, 3072, 256, 2), (64 * 41, 136, 768, 4)])
def test_gemm_wgrad_bias_side_product(dtype, T, N, K, split):
    """O2_EPI_ACCUM with trans_a and a bias pointer: dW += dY^T X AND db += colsum(dY) from one launch (tcgen05 arm: the
    epilogue warps add up the dY tiles of the items with n block 0 during the K loop).  Odd tile counts (the zero-row CTA of
    a pair), ragged T / N, several n blocks (K = 1024: 4 blocks of 256, so 3 of 4 items skip the sums), split-K tails, and
    accumulation INTO non-zero outputs."""
    from orbit2_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(T + N + split)
    dy = _mk((T, N), dtype, g); x = _mk((T, K), dtype, g)
    dw = torch.full((N, K), 0.5, device="cuda")
    db = torch.full((N,), -2.0, device="cuda")
    ops.gemm(dy, x, dw, trans_a=True, trans_b=True, epi=ops.EPI_ACCUM, split_k=split, bias=db)
    ref = dy.double().t() @ x.double() + 0.5
    refb = dy.double().sum(0) - 2.0
    tol = 2e-3 if dtype == torch.bfloat16 else 1e-5
    assert (dw - ref).abs().max().item() / ref.abs().max().item() < tol
    assert (db - refb).abs().max().item() / refb.abs().max().item() < 1e-5        # fp32 sums of the exact operand values
    # twice in a row on the same stream (ring phases carry over between items and launches start fresh)
    ops.gemm(dy, x, dw, trans_a=True, trans_b=True, epi=ops.EPI_ACCUM, split_k=split, bias=db)
    assert (db - (2 * refb + 2.0)).abs().max().item() / refb.abs().max().item() < 2e-5


@pytest.mark.parametrize("M,N,K,rps", [(300, 256, 128, 100), (1000, 1024, 256, 250), (257, 136, 64, 257)])
def test_gemm_fused_dropout_epilogues(M, N, K, rps):
    """o2_gemm_drop: the token-stream dropout / drop-path mask applied inside the BIAS_RES / BIAS_GELU / DGELU epilogues must
    be the SAME mask as the separate o2_dropout pass (oracle/dropout_mask.py) -- checked through the exact zero pattern --
    and the same values up to one bf16 rounding (the fused path rounds once, the two-pass path twice)."""
    from orbit2_b200 import ops
    from oracle.dropout_mask import keep_mask
    g = torch.Generator(device="cuda").manual_seed(M + N)
    bf = torch.bfloat16
    a = _mk((M, K), bf, g); w = _mk((N, K), bf, g); res = _mk((M, N), bf, g)
    bias = torch.randn(N, generator=g, device="cuda")
    ss = (torch.rand((M + rps - 1) // rps, generator=g, device="cuda") < 0.7).float() / 0.7
    p, seed = 0.25, 0x1234_5678_9ABC_DEF1
    keep = keep_mask(seed, 5, M * N, p).view(M, N).cuda()
    # BIAS_RES: x + drop_path(drop(proj(.)))
    fused = ops.gemm(a, w, torch.empty(M, N, device="cuda", dtype=bf), epi=ops.EPI_BIAS_RES, bias=bias, aux=res,
                     drop=(p, seed, 5, ss, rps))
    br = ops.gemm(a, w, torch.empty(M, N, device="cuda", dtype=bf), epi=ops.EPI_BIAS, bias=bias)
    two = ops.dropout(br, p, seed, 5, res=res, sample_scale=ss, rows_per_sample=rps)
    dead = (~keep) | (ss.repeat_interleave(rps)[:M, None] == 0)
    assert torch.equal(fused[dead], res[dead])
    assert rel(fused, two) < 1e-2
    exact = (a.double() @ w.double().t() + bias.double()) * keep.double() / (1 - p) * ss.repeat_interleave(rps)[:M, None].double() \
        + res.double()
    assert rel(fused, exact) < 6e-3
    # BIAS_GELU: drop1(gelu(fc1(.))), pre-activation kept unmasked
    pre = torch.empty(M, N, device="cuda", dtype=bf)
    h = ops.gemm(a, w, torch.empty(M, N, device="cuda", dtype=bf), epi=ops.EPI_BIAS_GELU, bias=bias, aux_out=pre,
                 drop=(p, seed, 5, None, 0))
    pre2 = torch.empty(M, N, device="cuda", dtype=bf)
    h2 = ops.gemm(a, w, torch.empty(M, N, device="cuda", dtype=bf), epi=ops.EPI_BIAS_GELU, bias=bias, aux_out=pre2)
    assert torch.equal(pre, pre2)
    assert torch.equal(h == 0, (~keep) | (h2 == 0))
    assert rel(h, ops.dropout(h2, p, seed, 5)) < 1e-2
    # DGELU: backward of the above
    dy = _mk((M, K), bf, g); wt = _mk((K, N), bf, g)
    d = ops.gemm(dy, wt, torch.empty(M, N, device="cuda", dtype=bf), trans_b=True, epi=ops.EPI_DGELU, aux=pre,
                 drop=(p, seed, 5, None, 0))
    d2 = ops.gemm(dy, wt, torch.empty(M, N, device="cuda", dtype=bf), trans_b=True, epi=ops.EPI_DGELU, aux=pre)
    assert torch.equal(d == 0, (~keep) | (d2 == 0))
    assert rel(d, ops.dropout(d2, p, seed, 5)) < 1e-2
