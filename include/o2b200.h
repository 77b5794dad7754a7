/* libo2b200 -- C ABI of the B200-native (sm_100a) kernels behind ORBIT-2's Reslim hot path.
 *
 * The reference (XiaoWang-Github/ORBIT-2) is 100% Python and has no FFI of its own: every device
 * kernel on this path is a PyTorch library call.  Each entry point below therefore names the
 * reference *call site* it replaces (paths relative to the reference root); INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers + sizes; no C++/torch types cross the boundary.
 *   - every pointer is a DEVICE pointer unless the name ends in _host; buffers are owned by the
 *     caller (torch allocator) and must outlive the launch; the library never allocates device memory.
 *   - `stream` is a cudaStream_t (as void*); all work is enqueued there, no internal synchronisation.
 *   - return 0 on success, a negative O2_ERR_* otherwise; o2_last_error() gives a thread-local message.
 *   - "act dtype" (O2_F32 / O2_BF16) is the activation storage type; all accumulation is fp32 (fp64
 *     for the loss sums).  Parameters that are reductions over tokens (d*-outputs marked "+=") are
 *     ACCUMULATED into, so the caller zero-fills them once per step.
 *   - no CPU fallback exists anywhere: without a CUDA device every compute call fails.
 */
#ifndef O2B200_H
#define O2B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define O2_OK 0
#define O2_ERR_ARG (-1)
#define O2_ERR_CUDA (-2)
#define O2_ERR_UNSUPPORTED (-3)

enum { O2_F32 = 0, O2_BF16 = 1 };

/* GEMM implementations: SIMT fp32 (exact-parity path, activations fp32) and tcgen05 bf16 (TMA + TMEM). */
enum { O2_GEMM_SIMT_F32 = 0, O2_GEMM_TC_BF16 = 1 };

/* GEMM epilogues (acc = op(A) op(B), fp32):                                                       */
enum {
  O2_EPI_NONE = 0,      /* C = acc                                                                */
  O2_EPI_BIAS = 1,      /* C = acc + bias[n]                                                      */
  O2_EPI_BIAS_GELU = 2, /* aux_out = acc + bias[n] (pre-activation, kept for backward); C = gelu   */
  O2_EPI_BIAS_RES = 3,  /* C = acc + bias[n] + aux[m % aux_rows, n] (residual / broadcast pos-emb) */
  O2_EPI_DGELU = 4,     /* C = acc * gelu'(aux[m, n])                                             */
  O2_EPI_ACCUM = 5      /* C(f32) += acc (split-K atomics; weight gradients).  With trans_a=1 a non-NULL
                         * `bias` is an OUTPUT: bias[m] (fp32 [M]) += sum_k A^T[m, k] -- the bias gradient
                         * db = colsum(dY) computed from the dY tiles the weight-gradient GEMM dW = dY^T X
                         * streams anyway (tcgen05 arm: by its epilogue warps during the K loop)        */
};

enum { O2_LOSS_MSE = 0, O2_LOSS_MAE = 1, O2_LOSS_BAYESIAN_TV = 2 };

int o2_version(void);
const char* o2_last_error(void);
/* 1 if a CUDA device of compute capability 10.x is visible, else 0 (never falls back). */
int o2_device_ok(void);

/* ---- dense contractions: every nn.Linear on the path -----------------------------------------
 * replaces F.linear in components/attention.py:50,81,177 (qkv/proj), components/mlp.py:63,67
 * (fc1/fc2), res_slimvit.py:115-120 (head) and their autograd backward.
 * C[M,N] = op(A)[M,K] * op(B)[K,N];  trans_a=0: A stored [M,K] row-major, 1: stored [K,M];
 * trans_b=0: B stored [N,K] row-major (nn.Linear weight layout), 1: stored [K,N].  ld* in elements.
 * impl SIMT: A,B,aux,aux_out fp32.  impl TC: A,B bf16 (16-byte aligned rows), aux/aux_out bf16,
 * C bf16 or fp32 (c_dtype).  bias is fp32 [N].  split_k>1 only with O2_EPI_ACCUM. */
int o2_gemm(int impl, const void* A, int trans_a, int64_t lda, const void* B, int trans_b, int64_t ldb, void* C,
            int c_dtype, int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const float* bias,
            const void* aux, int64_t ld_aux, int64_t aux_rows, void* aux_out, int64_t ld_aux_out, int split_k,
            void* stream);

/* ---- LayerNorm (vit_blocks.py:46,63 norm1/norm2; res_slimvit.py:104 norm), eps inside rsqrt ------- */
int o2_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int64_t T, int D, float eps, int dtype, void* stream);
/* dx = LN'(dy) (+ dres if non-null: the residual-stream gradient); dgamma/dbeta "+=" (fp32 [D]). */
int o2_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                     const void* dres, void* dx, float* dgamma, float* dbeta, int64_t T, int D, int dtype,
                     void* stream);

/* ---- multi-head self-attention, softmax(QK^T * scale) V, bidirectional, no mask -----------------
 * replaces attention.py:50-78 (xformers CK FMHA / F.scaled_dot_product_attention / explicit softmax).
 * qkv is the fused projection output [B, N, 3, heads, hd] (q rows, then k, then v: attention.py:50);
 * out is [B, N, heads, hd] (= x.transpose(1,2).reshape(B,N,C), attention.py:71,80); lse is
 * [B, heads, N] fp32 (natural log).  impl: O2_GEMM_SIMT_F32 (fp32 in/out) or O2_GEMM_TC_BF16. */
int o2_attn_fwd(int impl, const void* qkv, void* out, float* lse, int B, int N, int heads, int hd, float scale,
                void* stream);
/* dqkv has qkv's layout and is fully overwritten; delta is a [B,heads,N] fp32 scratch. */
int o2_attn_bwd(int impl, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                float* delta, int B, int N, int heads, int hd, float scale, void* stream);

/* The backward is three launches (delta = rowsum(dout*out); dK,dV; dQ).  o2_attn_bwd_parts runs the subset selected by
 * `parts` (bit mask) so that a caller can time them separately; O2_ATTN_BWD_DELTA must have run before the other two. */
enum { O2_ATTN_BWD_DELTA = 1, O2_ATTN_BWD_DKV = 2, O2_ATTN_BWD_DQ = 4, O2_ATTN_BWD_ALL = 7 };
int o2_attn_bwd_parts(int impl, int parts, const void* qkv, const void* out, const void* dout, const float* lse,
                      void* dqkv, float* delta, int B, int N, int heads, int hd, float scale, void* stream);

/* Training-mode variants with attention-probability dropout (attention.py:56,69,75: attn_drop on softmax(QK^T)): the output
 * uses P o keep / keep_prob, the softmax normaliser the unmasked P; the backward regenerates the same mask from
 * (seed, site) -- it is never stored.  keep(b, h, q, k) is one bit of a counter-based, bit-sliced 32-bit keep word per
 * (q, 32-key block) (csrc/common.cuh, restated in oracle/dropout_mask.py); p is quantised to floor(p * 65536) / 65536 (a
 * dithered byte threshold per keep word) and kept values are scaled by the exact keep probability.  p_drop = 0 is identical to the plain entry points. */
int o2_attn_fwd_drop(int impl, const void* qkv, void* out, float* lse, int B, int N, int heads, int hd, float scale,
                     float p_drop, uint64_t seed, uint32_t site, void* stream);
int o2_attn_bwd_parts_drop(int impl, int parts, const void* qkv, const void* out, const void* dout, const float* lse,
                           void* dqkv, float* delta, int B, int N, int heads, int hd, float scale, float p_drop,
                           uint64_t seed, uint32_t site, void* stream);

/* One-pass backward for head dim 64 on the tcgen05 arm (bf16): S = QK^T and dP = dO V^T are computed ONCE per (key tile,
 * query tile) pair -- 5 GEMMs and one softmax pass instead of the 7 + 2 of the two-kernel path above.  dK / dV accumulate
 * in tensor memory (single owner, exact); the dQ partial of every 128-key tile is added into `workspace` (fp32
 * [B, heads, N, hd], zero-filled by the call) by TMA reduce (L2 fp32 atomics: dQ is reproducible to fp32 rounding, not
 * bit for bit -- o2_attn_bwd_parts remains the deterministic option) and a last pass writes dQ = scale * workspace as
 * bf16.  Same dropout contract as o2_attn_bwd_parts_drop.  `parts` selects the launches so that a caller can time them:
 * O2_ATTN_BWD_DELTA, O2_ATTN_BWD_FUSED (workspace memset + the fused kernel), O2_ATTN_BWD_DQ_FINISH.
 * o2_attn_bwd_fused_workspace() = bytes the caller must provide (the reference has no counterpart: attention.py:54-78
 * leaves the backward to autograd / the FMHA library). */
enum { O2_ATTN_BWD_FUSED = 8, O2_ATTN_BWD_DQ_FINISH = 16, O2_ATTN_BWD_FUSED_ALL = 1 | 8 | 16 };
size_t o2_attn_bwd_fused_workspace(int B, int N, int heads, int hd);
int o2_attn_bwd_fused(int parts, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                      float* delta, void* workspace, size_t ws_bytes, int B, int N, int heads, int hd, float scale,
                      float p_drop, uint64_t seed, uint32_t site, void* stream);

/* ---- front end: per-variable patch embedding + variable embedding + variable aggregation ---------
 * replaces res_slimvit.py:254-265 (23x PatchEmbed conv, var_embed add, aggregate_variables) and
 * attention.py:132-176 (single-query cross attention over the V variable tokens) up to, not including,
 * var_agg.proj.  Uses the exact per-token collapse: the host precomputes
 *   tab_s [V, heads, PP+1]          score coefficients (last entry = constant term), pre-scaled
 *   tab_v [heads, V*(PP+1), hd]     value operator
 * (PP = patch_size^2) from the parameters; the kernel reads x [B, V, Hx, Wx] fp32 once and writes the
 * concatenated head outputs o [B, gh*gw, heads*hd] in act dtype.  Rows >= gh*p of x are ignored. */
int o2_frontend_fwd(const float* x, const float* tab_s, const float* tab_v, void* out, int out_dtype, int B, int V,
                    int Hx, int Wx, int p, int gh, int gw, int heads, int hd, void* stream);
/* dtab_s / dtab_v "+=" (fp32, same shapes as the tables). */
int o2_frontend_bwd(const float* x, const float* tab_s, const float* tab_v, const void* dout, int dtype,
                    float* dtab_s, float* dtab_v, int B, int V, int Hx, int Wx, int p, int gh, int gw, int heads,
                    int hd, void* stream);

/* ---- residual conv branch, first half (res_slimvit.py:107-109,233-242): gather the C+4 channels
 * ch_idx_host[0..cin) of x, conv3x3(cin -> c1, pad 1) + bias.  Writes the PRE-activation h1
 * [B, c1, Hx, Wx] in act dtype and, when g1 != NULL, also g1 = PixelShuffle(mag)(GELU(h1))
 * [B, c1/mag^2, Hx*mag, Wx*mag] (res_slimvit.py:109-110) for o2_headtail_* (g1 == NULL there: GELU +
 * PixelShuffle are applied on the fly from h1). */
int o2_path2_conv1_fwd(const float* x, const int* ch_idx_host, const float* w1, const float* b1, void* h1, void* g1,
                       int dtype, int B, int V, int Hx, int Wx, int cin, int c1, int mag, void* stream);
/* dw1 [c1,cin,3,3] / db1 [c1] "+=" from dh1 (gradient w.r.t. the pre-activation). */
int o2_path2_conv1_bwd(const float* x, const int* ch_idx_host, const void* dh1, float* dw1, float* db1, int dtype,
                       int B, int V, int Hx, int Wx, int cin, int c1, void* stream);

/* ---- head tail (res_slimvit.py:329-338): unpatchify (the reference's flat re-interpretation, see
 * SURVEY.md 8/a14) + conv_out 3x3 + [GELU -> PixelShuffle(mag) -> conv3x3(cr -> C)] of h1 + crop-add.
 * head_out [B, gh*gw, C*(mag*p)^2] act dtype; preds [B, C, Ho, Wo] act dtype, Ho = gh*p*mag, Wo = gw*p*mag;
 * h1 [B, cr*mag^2, Hx, Wx] with Hx*mag >= Ho, Wx*mag >= Wo (the branch is cropped to the ViT output);
 * g1 = the activated, shuffled branch written by o2_path2_conv1_fwd, or NULL. */
int o2_headtail_fwd(const void* head_out, const void* h1, const void* g1, const float* w_out, const float* b_out, const float* w2,
                    const float* b2, void* preds, int dtype, int B, int C, int gh, int gw, int p, int mag, int cr,
                    int Hx, int Wx, void* stream);
/* d_head_out / dh1 overwritten; dw_out [C,C,3,3], db_out [C], dw2 [C,cr,3,3], db2 [C] "+=". */
int o2_headtail_bwd(const void* dpreds, const void* head_out, const void* h1, const void* g1, const float* w_out,
                    const float* w2,
                    void* d_head_out, void* dh1, float* dw_out, float* db_out, float* dw2, float* db2, int dtype,
                    int B, int C, int gh, int gw, int p, int mag, int cr, int Hx, int Wx, void* stream);

/* ---- loss: clip_replace_constant (examples/intermediate_downscaling.py:267-278) fused with
 * mse / mae / bayesian_tv (metrics/functional.py:173-202, 218-232, 117-167) and their gradient.
 * pred [B,C,H,W] act dtype (raw model output, NOT modified); target fp32 [B,C,tgt_H,tgt_W] with
 * tgt_H>=H, tgt_W>=W (cropped like intermediate_downscaling.py:295-296).  clamp_ch: channel clamped
 * at 0 (-1: none); const_mask: bit c set => channel c replaced by the target (zero gradient).
 * lat_w fp32 [H] or NULL; ch_w fp32 [C] or NULL.  loss_vec [C+1] fp32: per-channel means then the
 * aggregate.  dpred (nullable) = grad_scale * d(aggregate)/d(pred), act dtype.  accum_ws: [C] fp64. */
int o2_loss_fwd_bwd(const void* pred, int dtype, const float* target, void* dpred, float* loss_vec,
                    double* accum_ws, const float* lat_w, const float* ch_w, int kind, int clamp_ch,
                    uint32_t const_mask, int B, int C, int H, int W, int tgt_H, int tgt_W, float grad_scale,
                    void* stream);
/* in-place clip_replace_constant for evaluation (intermediate_downscaling.py:333-334). */
int o2_clip_replace(void* pred, int dtype, const float* target, int clamp_ch, uint32_t const_mask, int B, int C,
                    int H, int W, int tgt_H, int tgt_W, void* stream);

/* g[b,c,:] *= scale[c] (device vector, fp32 [C]): chains an upstream gradient of the loss vector into dpred
 * (loss.backward() with a non-unit / per-channel grad_output; GradScaler-style loss scaling). */
int o2_scale_channels(void* g, int dtype, const float* scale, int B, int C, int64_t hw, void* stream);

/* ---- dropout / stochastic depth on the token stream ------------------------------------------
 * replaces nn.Dropout at res_slimvit.py:284 (pos_drop), components/attention.py:81 (proj_drop),
 * components/mlp.py:65,68 (drop1/drop2) and timm DropPath at components/vit_blocks.py:78-79, forward AND backward
 * (the same call on the gradient).  out[r,c] = res[r,c] + y[r,c] * keep(e)/(1-p) * sample_scale[r / rows_per_sample],
 * e = r*cols + c; res / sample_scale (fp32 [rows / rows_per_sample], per-sample drop-path factor) may be NULL;
 * out may alias y or res.  keep(e) is a counter-based hash of (seed, site, e) -- no mask is stored; the exact
 * function is documented in csrc/dropout.cu and restated in oracle/dropout_mask.py. */
int o2_dropout(const void* y, const void* res, void* out, int dtype, int64_t rows, int64_t cols,
               int64_t rows_per_sample, float p, const float* sample_scale, uint64_t seed, uint32_t site, void* stream);

/* CUDA-graph replay with dropout: after o2_dropout_seed_source(dev_word) every dropout-capable entry point called from this
 * host thread (o2_dropout, o2_gemm_drop, o2_attn_fwd_drop, o2_attn_bwd_parts_drop, o2_attn_bwd_fused) XORs a hash of the
 * 64-bit word at dev_word -- read ON THE DEVICE when the kernel runs -- into its seed, so a captured training step draws new
 * masks on every replay once the host rewrites the word in between (the by-value `seed` arguments are frozen into the
 * graph).  NULL restores by-value seeds.  Forward and backward kernels of one step must see the same word.  Restated in
 * oracle/dropout_mask.py (step_word argument). */
int o2_dropout_seed_source(const uint64_t* dev_word);

/* The same mask fused into the epilogue of the tcgen05 GEMM that produces the tensor (bf16 arm only; the fp32 arm keeps
 * the separate o2_dropout pass): e = m * N + n, mask m(e) = keep(e) / (1 - p), identical to o2_dropout on a [M, N] tensor.
 *   O2_EPI_BIAS_RES : C = (acc + bias) * m(e) * sample_scale[m / rows_per_sample] + aux     x + drop_path(proj_drop(proj(.)))
 *                     (after_residual: the mask multiplies the sum instead)
 *   O2_EPI_BIAS_GELU: aux_out = acc + bias;  C = gelu(aux_out) * m(e)                        drop1(act(fc1(.)))  mlp.py:63-65
 *   O2_EPI_DGELU    : C = acc * m(e) * gelu'(aux)                                            backward of the line above
 * sample_scale may be NULL (no drop-path); p may be 0 (drop-path only). */
typedef struct {
  float p;
  uint64_t seed;
  uint32_t site;
  const float* sample_scale; /* fp32 [M / rows_per_sample] or NULL */
  int64_t rows_per_sample;
  int32_t after_residual;    /* O2_EPI_BIAS_RES only: C = (acc + bias + aux) * m(e) -- pos_drop(tokens + pos_embed),
                              * res_slimvit.py:281-284 -- instead of (acc + bias) * m(e) + aux */
} O2GemmDrop;
/* LayerNorm backward with a second, masked output (bf16, D <= 1024): dx as o2_layernorm_bwd, and dx_masked = dx * m(e) *
 * sample_scale[row / rows_per_sample], e = row * D + col -- the gradient entering a residual branch whose forward was
 * x + drop_path(drop(branch(.))), without a separate pass over dx. */
int o2_layernorm_bwd_drop(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                          const void* dres, void* dx, void* dx_masked, float* dgamma, float* dbeta, int64_t T, int D,
                          const O2GemmDrop* drop, void* stream);
int o2_gemm_drop(const void* A, int trans_a, int64_t lda, const void* B, int trans_b, int64_t ldb, void* C, int64_t ldc,
                 int64_t M, int64_t N, int64_t K, int epilogue, const float* bias, const void* aux, int64_t ld_aux,
                 int64_t aux_rows, void* aux_out, int64_t ld_aux_out, const O2GemmDrop* drop, void* stream);

/* ---- either side of the hot path: input normalisation and evaluation statistics ---------------
 * o2_normalize_fields replaces the per-sample host transforms of the data pipeline (data/itermodule.py:202-211,
 * iterdataset.py:360-379): x [B,V,hw] fp32 RAW fields, in place; kind[v] = 0: (x - mean[v]) / std[v] (torchvision
 * Normalize), 1: LogTransform (precipmodule.py:21-42: x*1000, values <= 0.25 -> 0, log1p).  mean/std/kind: device [V].
 * o2_eval_stats feeds rmse / pearson / mean_bias (metrics/functional.py:236-257, 294-324): out [B,C,6] fp64 (zeroed by
 * the call) = sums over H x W of { w e^2, p, t, p^2, t^2, p t } with p = scale[c]*pred + shift[c], t likewise (the
 * denormalising TransformedMetric, metrics/metrics.py:100-115; NULL = identity), e = p - t, w = lat_w[y] or 1. */
int o2_normalize_fields(float* x, const float* mean, const float* stdv, const int* kind, int B, int V, int64_t hw,
                        void* stream);
int o2_eval_stats(const void* pred, int dtype, const float* target, const float* lat_w, const float* scale,
                  const float* shift, double* out, int B, int C, int H, int W, int tgt_H, int tgt_W, void* stream);

/* ---- small HBM-bound helpers ---------------------------------------------------------------- */
int o2_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
/* bf16 -> fp32: operands of the fp32 attention arm for head dims the tcgen05 kernels do not cover (256, interm_10b). */
int o2_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream);
/* ---- position-embedding resample ------------------------------------------------------------------
 * replaces interpolate_pos_embed_on_the_fly (components/pos_embed.py:103-138, called every forward at
 * res_slimvit.py:271-278 when the token grid differs from the stored one: TILES mode, other-resolution inference):
 * torch's bicubic upsample (align_corners=False, A = -0.75, taps clamped to the border) on a channels-last
 * [ih*iw, D] fp32 table -> [oh*ow, D], and its exact adjoint (d_src OVERWRITTEN; gather form, deterministic). */
int o2_bicubic_fwd(const float* src, float* dst, int ih, int iw, int oh, int ow, int D, void* stream);
int o2_bicubic_bwd(const float* d_dst, float* d_src, int ih, int iw, int oh, int ow, int D, void* stream);

/* out[n] += sum_m X[m, n]  (bias gradients), X act dtype with row pitch ld. */
int o2_colsum(const void* X, int dtype, float* out, int64_t M, int64_t N, int64_t ld, void* stream);
/* fused AdamW (torch.optim.AdamW semantics, intermediate_downscaling.py:642-644) over a flat fp32
 * buffer; g is multiplied by grad_scale first; optionally refreshes the bf16 compute copy. */
int o2_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
             float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);
/* the same update with the step-dependent scalars read from DEVICE memory, so that a captured CUDA graph of the whole
 * training step can be replayed while the host rewrites them: scalars fp32 [8] = {lr, beta1, beta2, eps, weight_decay,
 * 1 - beta1^step, sqrt(1 - beta2^step), grad_scale}. */
int o2_adamw_dev(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, const float* scalars,
                 void* stream);
/* *flag |= 1 if any of g[0..n) is inf or NaN: the found_inf test of the reference's bf16 branch
 * (ShardedGradScaler.step, examples/intermediate_downscaling.py:733-742).  The caller zero-fills flag. */
int o2_nonfinite(const float* g, int64_t n, int* flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif
