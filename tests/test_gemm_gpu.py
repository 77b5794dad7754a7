"""GPU parity: tcgen05 bf16 GEMM and SIMT fp32 GEMM vs torch fp32 matmul of the same (bf16-rounded) inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


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
