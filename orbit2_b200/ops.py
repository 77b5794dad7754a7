"""Thin tensor -> pointer wrappers over the C ABI.  Every function launches on torch's current stream
and raises O2Error on failure.  No op has a PyTorch/CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib as L
from ._lib import (EPI_ACCUM, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_DGELU, EPI_NONE, GEMM_SIMT_F32, GEMM_TC_BF16,
                   O2_BF16, O2_F32)

TIMERS = None         # bench.py sets this to {} to collect (name -> [(start_event, end_event), ...]) per kernel
LAUNCHES = 0          # number of library kernels-launching calls (bench.py reports it)


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda, "orbit2_b200 ops need CUDA tensors (no CPU fallback)"
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return O2_F32
    if t.dtype == torch.bfloat16:
        return O2_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def impl_for(dtype: torch.dtype) -> int:
    return GEMM_TC_BF16 if dtype == torch.bfloat16 else GEMM_SIMT_F32


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, trans_a=False, trans_b=False, epi=EPI_NONE,
         bias=None, aux=None, aux_rows=0, aux_out=None, split_k=1, M=None, N=None, K=None, drop=None):
    """out[M,N] = op(a) @ op(b) with fused epilogue; a/b/out are 2-D row-major (last stride 1).
    ``drop`` = (p, seed, site, sample_scale or None, rows_per_sample[, after_residual]): the token-stream dropout / drop-path
    mask of o2_dropout fused into the BIAS_RES / BIAS_GELU / DGELU epilogue (bf16 arm only, o2_gemm_drop)."""
    lib = L.load()
    assert a.dim() == 2 and b.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    if M is None:
        M = a.shape[1] if trans_a else a.shape[0]
    if K is None:
        K = a.shape[0] if trans_a else a.shape[1]
    if N is None:
        N = b.shape[1] if trans_b else b.shape[0]
    assert (b.shape[0] if trans_b else b.shape[1]) == K, (a.shape, b.shape, trans_a, trans_b)
    assert out.shape[0] == M and out.shape[1] == N
    impl = impl_for(a.dtype)
    assert b.dtype == a.dtype
    if drop is not None:
        assert impl == GEMM_TC_BF16 and split_k == 1 and out.dtype == torch.bfloat16, "fused dropout: bf16 arm only"
        p, seed, site, ss, rps = drop[:5]
        spec = L.GemmDrop(float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(site) & 0xFFFFFFFF, ss.data_ptr() if ss is not None else None,
                          int(rps) if ss is not None else 0, 1 if (len(drop) > 5 and drop[5]) else 0)
        with _timed("gemm"):
            rc = lib.o2_gemm_drop(_ptr(a), int(trans_a), a.stride(0), _ptr(b), int(trans_b), b.stride(0), _ptr(out),
                                  out.stride(0), M, N, K, epi, _ptr(bias), _ptr(aux), aux.stride(0) if aux is not None else 0,
                                  aux_rows, _ptr(aux_out), aux_out.stride(0) if aux_out is not None else 0, C.byref(spec),
                                  _stream())
        if TIMERS is not None:
            TIMERS.setdefault("gemm_flops", []).append(2.0 * M * N * K)
        L.check(rc, "o2_gemm_drop")
        _count()
        return out
    with _timed("gemm"):
        rc = lib.o2_gemm(impl, _ptr(a), int(trans_a), a.stride(0), _ptr(b), int(trans_b), b.stride(0), _ptr(out), dt(out),
                         out.stride(0), M, N, K, epi, _ptr(bias), _ptr(aux), aux.stride(0) if aux is not None else 0,
                         aux_rows, _ptr(aux_out), aux_out.stride(0) if aux_out is not None else 0, split_k, _stream())
    if TIMERS is not None:
        TIMERS.setdefault("gemm_flops", []).append(2.0 * M * N * K)
    L.check(rc, "o2_gemm")
    _count()
    return out


def layernorm_fwd(x, gamma, beta, eps=1e-5):
    """x [T,D] act dtype; gamma/beta fp32 -> (y, mean, rstd)."""
    lib = L.load()
    T, D = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(T, device=x.device, dtype=torch.float32)
    rstd = torch.empty(T, device=x.device, dtype=torch.float32)
    L.check(lib.o2_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean), _ptr(rstd), T, D, eps, dt(x),
                                 _stream()), "o2_layernorm_fwd")
    _count()
    return y, mean, rstd


LN_DROP_MAX_D = 1024      # widest row of the bf16 LayerNorm-backward kernel that can emit the masked second output


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dres=None, drop=None):
    """returns dx (= LN'(dy) + dres); dgamma/dbeta fp32 are accumulated into.  With ``drop`` = (p, seed, site, sample_scale or
    None, rows_per_sample) returns (dx, dx * mask): the second tensor is what o2_dropout would make of dx (bf16, D <= 1024:
    written by the same kernel; otherwise by a separate o2_dropout pass)."""
    lib = L.load()
    T, D = x.shape
    dx = torch.empty_like(x)
    if drop is not None and x.dtype == torch.bfloat16 and D <= LN_DROP_MAX_D and not os.environ.get("O2_LN_BWD_SPLIT"):
        p, seed, site, ss, rps = drop
        spec = L.GemmDrop(float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(site) & 0xFFFFFFFF, ss.data_ptr() if ss is not None else None,
                          int(rps) if ss is not None else 0, 0)
        dxm = torch.empty_like(x)
        L.check(lib.o2_layernorm_bwd_drop(_ptr(dy), _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dres), _ptr(dx), _ptr(dxm),
                                          _ptr(dgamma), _ptr(dbeta), T, D, C.byref(spec), _stream()), "o2_layernorm_bwd_drop")
        _count()
        return dx, dxm
    L.check(lib.o2_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dres), _ptr(dx),
                                 _ptr(dgamma), _ptr(dbeta), T, D, dt(x), _stream()), "o2_layernorm_bwd")
    _count()
    if drop is not None:
        p, seed, site, ss, rps = drop
        return dx, dropout(dx, p, seed, site, sample_scale=ss, rows_per_sample=rps if ss is not None else 0)
    return dx


class _timed:
    """Records a CUDA-event pair around a launch on the current stream when TIMERS is enabled."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMERS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if TIMERS is not None:
            self.e1.record()
            TIMERS.setdefault(self.name, []).append((self.e0, self.e1))
        return False


def cast_f32(src: torch.Tensor):
    """bf16 -> fp32 copy (o2_cast_bf16_to_f32)."""
    lib = L.load()
    dst = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    L.check(lib.o2_cast_bf16_to_f32(_ptr(src), _ptr(dst), src.numel(), _stream()), "o2_cast_bf16_to_f32")
    _count()
    return dst


# head dims of the tcgen05 attention kernels (256 = interm_10b: csrc/attn_tc256.cu, no attention-dropout variant); a bf16
# call outside this set runs the fp32 SIMT arm on up-cast operands and rounds the result back
TC_HEAD_DIMS = (64, 128, 256)
TC_NO_DROPOUT_HEAD_DIMS = (256,)


def _tc_ok(hd, drop):
    return hd in TC_HEAD_DIMS and not (hd in TC_NO_DROPOUT_HEAD_DIMS and drop is not None and drop[0] > 0.0)


def attn_fwd(qkv, B, N, heads, hd, drop=None):
    """qkv [B*N, 3*heads*hd] -> (out [B*N, heads*hd], lse [B,heads,N]).  drop = (p, seed, site): attention-probability
    dropout (training mode)."""
    lib = L.load()
    if qkv.dtype == torch.bfloat16 and not _tc_ok(hd, drop):
        out32, lse = attn_fwd(cast_f32(qkv), B, N, heads, hd, drop)
        return cast_bf16(out32), lse
    out = torch.empty(B * N, heads * hd, device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty(B, heads, N, device=qkv.device, dtype=torch.float32)
    p, seed, site = drop if drop is not None else (0.0, 0, 0)
    with _timed("attn_fwd"):
        L.check(lib.o2_attn_fwd_drop(impl_for(qkv.dtype), _ptr(qkv), _ptr(out), _ptr(lse), B, N, heads, hd, hd ** -0.5,
                                     float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(site) & 0xFFFFFFFF, _stream()),
                "o2_attn_fwd")
    _count()
    return out, lse


# head dims of the one-pass backward (5 GEMMs, dQ partials reduced through the L2 by TMA); O2_ATTN_BWD_TWO_PASS=1 keeps the
# deterministic two-kernel backward (7 GEMMs, no atomics)
FUSED_BWD_HEAD_DIMS = (64,)


def attn_bwd_two_pass_default() -> bool:
    return bool(int(os.environ.get("O2_ATTN_BWD_TWO_PASS", "0")))         # read per call: tests flip it around engine runs


def attn_bwd(qkv, out, dout, lse, B, N, heads, hd, drop=None, two_pass=None):
    """-> dqkv.  bf16, head dim 64: the one-pass kernel unless ``two_pass`` (default: O2_ATTN_BWD_TWO_PASS)."""
    lib = L.load()
    if qkv.dtype == torch.bfloat16 and not _tc_ok(hd, drop):
        return cast_bf16(attn_bwd(cast_f32(qkv), cast_f32(out), cast_f32(dout), lse, B, N, heads, hd, drop))
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, heads, N, device=qkv.device, dtype=torch.float32)
    impl = impl_for(qkv.dtype)
    p, seed, site = drop if drop is not None else (0.0, 0, 0)
    if two_pass is None:
        two_pass = attn_bwd_two_pass_default()
    if impl == GEMM_TC_BF16 and hd in FUSED_BWD_HEAD_DIMS and not two_pass:
        ws_bytes = int(lib.o2_attn_bwd_fused_workspace(B, N, heads, hd))
        ws = torch.empty(ws_bytes // 4, device=qkv.device, dtype=torch.float32)        # caller-owned scratch (dQ accumulator)
        fargs = (_ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv), _ptr(delta), _ptr(ws), ws_bytes, B, N, heads, hd,
                 hd ** -0.5, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(site) & 0xFFFFFFFF, _stream())
        if TIMERS is not None:
            for name, part in (("attn_bwd_delta", 1), ("attn_bwd_fused", 8), ("attn_bwd_dq_finish", 16)):
                with _timed(name):
                    L.check(lib.o2_attn_bwd_fused(part, *fargs), "o2_attn_bwd_fused")
        else:
            L.check(lib.o2_attn_bwd_fused(1 | 8 | 16, *fargs), "o2_attn_bwd_fused")
        _count(4)
        return dqkv
    args = (_ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv), _ptr(delta), B, N, heads, hd, hd ** -0.5, float(p),
            int(seed) & 0xFFFFFFFFFFFFFFFF, int(site) & 0xFFFFFFFF, _stream())
    if TIMERS is not None and impl == GEMM_TC_BF16:
        for name, part in (("attn_bwd_delta", 1), ("attn_bwd_dkv", 2), ("attn_bwd_dq", 4)):
            with _timed(name):
                L.check(lib.o2_attn_bwd_parts_drop(impl, part, *args), "o2_attn_bwd_parts")
    else:
        L.check(lib.o2_attn_bwd_parts_drop(impl, 7, *args), "o2_attn_bwd")
    _count(3)
    return dqkv


def cast_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None):
    lib = L.load()
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    L.check(lib.o2_cast_f32_to_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()), "o2_cast_f32_to_bf16")
    _count()
    return dst


def bicubic_fwd(src: torch.Tensor, ih: int, iw: int, oh: int, ow: int) -> torch.Tensor:
    """[ih*iw, D] fp32 -> [oh*ow, D]: torch's bicubic (align_corners=False) on a channels-last table (pos_embed.py:103-138)."""
    lib = L.load()
    assert src.dtype == torch.float32 and src.is_contiguous() and src.shape[0] == ih * iw
    dst = torch.empty(oh * ow, src.shape[1], device=src.device, dtype=torch.float32)
    L.check(lib.o2_bicubic_fwd(_ptr(src), _ptr(dst), ih, iw, oh, ow, src.shape[1], _stream()), "o2_bicubic_fwd")
    _count()
    return dst


def bicubic_bwd(d_dst: torch.Tensor, ih: int, iw: int, oh: int, ow: int) -> torch.Tensor:
    """adjoint of bicubic_fwd: [oh*ow, D] -> [ih*iw, D] (deterministic gather)."""
    lib = L.load()
    assert d_dst.dtype == torch.float32 and d_dst.is_contiguous() and d_dst.shape[0] == oh * ow
    d_src = torch.empty(ih * iw, d_dst.shape[1], device=d_dst.device, dtype=torch.float32)
    L.check(lib.o2_bicubic_bwd(_ptr(d_dst), _ptr(d_src), ih, iw, oh, ow, d_dst.shape[1], _stream()), "o2_bicubic_bwd")
    _count()
    return d_src


def colsum(x: torch.Tensor, out: torch.Tensor):
    """out[N] += sum over rows of x [M,N]."""
    lib = L.load()
    L.check(lib.o2_colsum(_ptr(x), dt(x), _ptr(out), x.shape[0], x.shape[1], x.stride(0), _stream()), "o2_colsum")
    _count()
    return out


def adamw(p, g, m, v, p_bf16, lr, beta1, beta2, eps, wd, step, grad_scale=1.0):
    lib = L.load()
    L.check(lib.o2_adamw(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_bf16), p.numel(), lr, beta1, beta2, eps, wd, step,
                         grad_scale, _stream()), "o2_adamw")
    _count()


def nonfinite(g, flag):
    """flag (int32 [1], device, zero-filled by the caller) |= 1 if any element of the fp32 tensor g is inf / NaN."""
    lib = L.load()
    assert g.dtype == torch.float32 and g.is_contiguous() and flag.dtype == torch.int32
    L.check(lib.o2_nonfinite(_ptr(g), g.numel(), _ptr(flag), _stream()), "o2_nonfinite")
    _count()


def adamw_dev(p, g, m, v, p_bf16, scalars):
    """AdamW with {lr, beta1, beta2, eps, wd, 1 - beta1^t, sqrt(1 - beta2^t), grad_scale} read from the device tensor
    ``scalars`` (fp32 [8]) -- the form a captured CUDA graph of the step replays."""
    lib = L.load()
    assert scalars.is_cuda and scalars.dtype == torch.float32 and scalars.numel() >= 8
    L.check(lib.o2_adamw_dev(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_bf16), p.numel(), _ptr(scalars), _stream()),
            "o2_adamw_dev")
    _count()


def loss_fwd_bwd(pred, target, kind, *, lat_w=None, ch_w=None, clamp_ch=-1, const_mask=0, want_grad=True,
                 grad_scale=1.0):
    """pred [B,C,H,W] act dtype (raw); target fp32 [B,C,tH,tW].  Returns (loss_vec [C+1] fp32, dpred or None)."""
    lib = L.load()
    B, Cc, H, W = pred.shape
    tH, tW = target.shape[2], target.shape[3]
    assert pred.is_contiguous() and target.is_contiguous() and target.dtype == torch.float32
    loss_vec = torch.empty(Cc + 1, device=pred.device, dtype=torch.float32)
    ws = torch.empty(Cc, device=pred.device, dtype=torch.float64)
    dpred = torch.empty_like(pred) if want_grad else None
    L.check(lib.o2_loss_fwd_bwd(_ptr(pred), dt(pred), _ptr(target), _ptr(dpred), _ptr(loss_vec), _ptr(ws), _ptr(lat_w),
                                _ptr(ch_w), kind, clamp_ch, const_mask, B, Cc, H, W, tH, tW, grad_scale, _stream()),
            "o2_loss_fwd_bwd")
    _count(2)
    return loss_vec, dpred


def clip_replace_(pred, target, clamp_ch, const_mask):
    lib = L.load()
    B, Cc, H, W = pred.shape
    L.check(lib.o2_clip_replace(_ptr(pred), dt(pred), _ptr(target), clamp_ch, const_mask, B, Cc, H, W, target.shape[2],
                                target.shape[3], _stream()), "o2_clip_replace")
    _count()
    return pred


def _heads_of(tab_s):
    return tab_s.shape[1]


FE_MAX_HD = 128


def _split_wide_heads(tab_s, tab_v, hd):
    """Head dims above the front-end kernels' 128: the softmax over the variables depends on the head, the output columns
    are independent, so a head of width hd is f = hd / 128 heads of width 128 that share one score table (tab_s rows
    repeated, tab_v [heads, KK, hd] re-laid as [heads * f, KK, 128]); the output [T, heads * hd] is the same memory."""
    f = hd // FE_MAX_HD
    assert hd % FE_MAX_HD == 0
    heads, KK = tab_v.shape[0], tab_v.shape[1]
    ts = tab_s.repeat_interleave(f, dim=1).contiguous()
    tv = tab_v.view(heads, KK, f, FE_MAX_HD).permute(0, 2, 1, 3).reshape(heads * f, KK, FE_MAX_HD).contiguous()
    return ts, tv, f


def frontend_fwd(x, tab_s, tab_v, p, gh, gw, hd, out_dtype):
    """x [B,V,Hx,Wx] fp32; tab_s [V,heads,PP+1]; tab_v [heads, V*(PP+1), hd] -> o [B*gh*gw, heads*hd]."""
    lib = L.load()
    if hd > FE_MAX_HD:
        ts, tv, _ = _split_wide_heads(tab_s, tab_v, hd)
        return frontend_fwd(x, ts, tv, p, gh, gw, FE_MAX_HD, out_dtype)
    B, V, Hx, Wx = x.shape
    heads = _heads_of(tab_s)
    assert x.dtype == torch.float32 and x.is_contiguous() and tab_s.is_contiguous() and tab_v.is_contiguous()
    out = torch.empty(B * gh * gw, heads * hd, device=x.device, dtype=out_dtype)
    L.check(lib.o2_frontend_fwd(_ptr(x), _ptr(tab_s), _ptr(tab_v), _ptr(out), dt(out), B, V, Hx, Wx, p, gh, gw, heads, hd,
                                _stream()), "o2_frontend_fwd")
    _count()
    return out


def frontend_bwd(x, tab_s, tab_v, dout, p, gh, gw, hd):
    """-> (dtab_s, dtab_v) fp32."""
    lib = L.load()
    if hd > FE_MAX_HD:
        ts, tv, f = _split_wide_heads(tab_s, tab_v, hd)
        dts, dtv = frontend_bwd(x, ts, tv, dout, p, gh, gw, FE_MAX_HD)
        heads, KK = tab_v.shape[0], tab_v.shape[1]
        dts = dts.view(tab_s.shape[0], heads, f, tab_s.shape[2]).sum(2)
        dtv = dtv.view(heads, f, KK, FE_MAX_HD).permute(0, 2, 1, 3).reshape(heads, KK, hd)
        return dts.contiguous(), dtv.contiguous()
    B, V, Hx, Wx = x.shape
    heads = _heads_of(tab_s)
    dts = torch.zeros_like(tab_s)
    dtv = torch.zeros_like(tab_v)
    L.check(lib.o2_frontend_bwd(_ptr(x), _ptr(tab_s), _ptr(tab_v), _ptr(dout), dt(dout), _ptr(dts), _ptr(dtv), B, V, Hx,
                                Wx, p, gh, gw, heads, hd, _stream()), "o2_frontend_bwd")
    _count()
    return dts, dtv


def _idx(ch_idx):
    return (C.c_int * len(ch_idx))(*[int(i) for i in ch_idx])


def path2_conv1_fwd(x, ch_idx, w1, b1, out_dtype, mag=0):
    """x [B,V,Hx,Wx] fp32 -> pre-activation h1 [B, c1, Hx, Wx]; with mag > 0 also g1 = PixelShuffle(mag)(GELU(h1))
    [B, c1/mag^2, Hx*mag, Wx*mag] and the return value is (h1, g1)."""
    lib = L.load()
    B, V, Hx, Wx = x.shape
    c1, cin = w1.shape[0], w1.shape[1]
    assert cin == len(ch_idx)
    h1 = torch.empty(B, c1, Hx, Wx, device=x.device, dtype=out_dtype)
    g1 = torch.empty(B, c1 // (mag * mag), Hx * mag, Wx * mag, device=x.device, dtype=out_dtype) if mag else None
    L.check(lib.o2_path2_conv1_fwd(_ptr(x), _idx(ch_idx), _ptr(w1), _ptr(b1), _ptr(h1), _ptr(g1) if mag else None, dt(h1),
                                   B, V, Hx, Wx, cin, c1, mag, _stream()), "o2_path2_conv1_fwd")
    _count()
    return (h1, g1) if mag else h1


def path2_conv1_bwd(x, ch_idx, dh1, dw1, db1):
    """dw1/db1 fp32 are accumulated into."""
    lib = L.load()
    B, V, Hx, Wx = x.shape
    c1, cin = dw1.shape[0], dw1.shape[1]
    L.check(lib.o2_path2_conv1_bwd(_ptr(x), _idx(ch_idx), _ptr(dh1), _ptr(dw1), _ptr(db1), dt(dh1), B, V, Hx, Wx, cin, c1,
                                   _stream()), "o2_path2_conv1_bwd")
    _count()


def headtail_fwd(head_out, h1, w_out, b_out, w2, b2, B, C_, gh, gw, p, mag, g1=None):
    """head_out [B*gh*gw, C*(mag*p)^2], h1 [B, cr*mag^2, Hx, Wx] -> preds [B, C, gh*p*mag, gw*p*mag]."""
    lib = L.load()
    cr = w2.shape[1]
    Hx, Wx = h1.shape[2], h1.shape[3]
    preds = torch.empty(B, C_, gh * p * mag, gw * p * mag, device=head_out.device, dtype=head_out.dtype)
    L.check(lib.o2_headtail_fwd(_ptr(head_out), _ptr(h1), _ptr(g1) if g1 is not None else None, _ptr(w_out), _ptr(b_out), _ptr(w2), _ptr(b2), _ptr(preds),
                                dt(preds), B, C_, gh, gw, p, mag, cr, Hx, Wx, _stream()), "o2_headtail_fwd")
    _count()
    return preds


def headtail_bwd(dpreds, head_out, h1, w_out, w2, dw_out, db_out, dw2, db2, B, C_, gh, gw, p, mag, g1=None):
    """-> (d_head_out, dh1); the four weight gradients (fp32) are accumulated into."""
    lib = L.load()
    cr = w2.shape[1]
    Hx, Wx = h1.shape[2], h1.shape[3]
    dho = torch.empty_like(head_out)
    dh1 = torch.empty_like(h1)
    assert dpreds.is_contiguous() and dpreds.dtype == head_out.dtype
    L.check(lib.o2_headtail_bwd(_ptr(dpreds), _ptr(head_out), _ptr(h1), _ptr(g1) if g1 is not None else None, _ptr(w_out),
                                _ptr(w2), _ptr(dho), _ptr(dh1),
                                _ptr(dw_out), _ptr(db_out), _ptr(dw2), _ptr(db2), dt(dpreds), B, C_, gh, gw, p, mag, cr,
                                Hx, Wx, _stream()), "o2_headtail_bwd")
    _count()
    return dho, dh1


def scale_channels_(g, scale):
    """g [B,C,H,W] *= scale[c] (fp32 device vector), in place."""
    lib = L.load()
    B, Cc, H, W = g.shape
    L.check(lib.o2_scale_channels(_ptr(g), dt(g), _ptr(scale), B, Cc, H * W, _stream()), "o2_scale_channels")
    _count()
    return g


def dropout(y, p, seed, site, *, res=None, sample_scale=None, rows_per_sample=0, out=None):
    """out = res + y * keep(seed, site, element) / (1 - p) * sample_scale[row // rows_per_sample]; y [rows, cols].
    The same call on a gradient is the backward.  ``out`` may be ``y`` (in place)."""
    lib = L.load()
    rows, cols = y.shape
    assert y.is_contiguous() and (res is None or (res.is_contiguous() and res.shape == y.shape and res.dtype == y.dtype))
    if out is None:
        out = torch.empty_like(y)
    L.check(lib.o2_dropout(_ptr(y), _ptr(res), _ptr(out), dt(y), rows, cols, rows_per_sample, float(p), _ptr(sample_scale),
                           int(seed) & 0xFFFFFFFFFFFFFFFF, int(site) & 0xFFFFFFFF, _stream()), "o2_dropout")
    _count()
    return out


def dropout_seed_source(word: Optional[torch.Tensor]):
    """Install (or with None remove) the device-resident 64-bit step word that every dropout-capable call of this thread XORs
    into its seed at kernel run time (include/o2b200.h o2_dropout_seed_source): masks that change between CUDA-graph replays."""
    lib = L.load()
    if word is not None:
        assert word.is_cuda and word.dtype == torch.int64 and word.numel() == 1
    L.check(lib.o2_dropout_seed_source(_ptr(word)), "o2_dropout_seed_source")


def normalize_fields_(x, mean, std, kind):
    """x [B,V,H,W] fp32 raw fields, normalised in place; mean/std fp32 [V], kind int32 [V] (0 Normalize, 1 LogTransform)."""
    lib = L.load()
    B, V = x.shape[0], x.shape[1]
    assert x.dtype == torch.float32 and x.is_contiguous() and kind.dtype == torch.int32
    L.check(lib.o2_normalize_fields(_ptr(x), _ptr(mean), _ptr(std), _ptr(kind), B, V, x[0, 0].numel(), _stream()),
            "o2_normalize_fields")
    _count()
    return x


def eval_stats(pred, target, *, lat_w=None, scale=None, shift=None):
    """-> [B, C, 6] fp64 sums {w e^2, p, t, p^2, t^2, p t} (see o2b200.h)."""
    lib = L.load()
    B, Cc, H, W = pred.shape
    assert pred.is_contiguous() and target.is_contiguous() and target.dtype == torch.float32
    out = torch.empty(B, Cc, 6, device=pred.device, dtype=torch.float64)
    L.check(lib.o2_eval_stats(_ptr(pred), dt(pred), _ptr(target), _ptr(lat_w), _ptr(scale), _ptr(shift), _ptr(out), B, Cc,
                              H, W, target.shape[2], target.shape[3], _stream()), "o2_eval_stats")
    _count(2)
    return out
