// Front end: per-variable patch embedding + variable embedding + single-query cross-attention variable aggregation,
// collapsed EXACTLY into a per-token function of the raw p x p pixel patches.
//
// reference (as written): res_slimvit.py:254-265 -- V x Conv2d(1->D,k=p,s=p), stack to [B,V,L,D] (763 MB/sample at 117M),
// + var_embed, permute to [B*L,V,D]; attention.py:142-176 -- kv = Linear(D,2D) over all B*L*V tokens (1.56 TFLOP/sample),
// softmax over the V keys against ONE learned query, 16 heads.  Because the query is token-independent and every key/value
// token is affine in the patch pixels, with P' = [pixels, 1]:
//     score[t,v,h] = tab_s[v,h,:] . P'[t,v,:]            a = softmax_v(score)
//     o[t,h,:]     = sum_{v,k'} (a[t,v,h] P'[t,v,k']) tab_v[h,(v,k'),:]
// The host builds tab_s / tab_v from the parameters (tiny, autograd-visible); this kernel reads x once per head
// (x stays L2-resident) and writes o [B*L, heads*hd]; var_agg.proj follows as a tensor-core GEMM.
// Backward reduces d(tab_s), d(tab_v) over all tokens: persistent CTAs keep their slice in registers, one atomic
// per element per CTA at the end.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TT = 64;      // tokens per tile
constexpr int NT = 256;
constexpr int KPT = 8;      // table rows per thread in the d(tab_v) accumulation -> KK <= 128

struct FeArgs {
  const float* x; const float* tab_s; const float* tab_v;
  void* out; const void* dout; float* dtab_s; float* dtab_v;
  int B, V, Hx, Wx, p, gh, gw, heads, hd, PP, KK;
  int VP, fast;     // bf16 arm, p = 2: coefficient columns / table rows in (k', v) order, kk' = k' * VP + v (VP = V rounded up to even)
  long long T;
};

// stage pixels P[t][v][k] and softmax weights a[t][v] of one token tile for head h
__device__ __forceinline__ void stage_tile(const FeArgs& a, int h, long long t0, float* sP, float* sa, float* ssc) {
  const int V = a.V, PP = a.PP, P1 = PP + 1;
  const int L = a.gh * a.gw;
  for (int i = threadIdx.x; i < TT * V; i += NT) {
    const int tl = i % TT, v = i / TT;
    const long long t = t0 + tl;
    float sc = -INFINITY;
    if (t < a.T) {
      const int b = (int)(t / L), l = (int)(t % L);
      const int gy = l / a.gw, gx = l % a.gw;
      const float* xp = a.x + (((size_t)b * V + v) * a.Hx + (size_t)gy * a.p) * a.Wx + (size_t)gx * a.p;
      const float* ts = a.tab_s + ((size_t)v * a.heads + h) * P1;
      sc = ts[PP];
      for (int pi = 0; pi < a.p; ++pi)
        for (int pj = 0; pj < a.p; ++pj) {
          const float px = xp[(size_t)pi * a.Wx + pj];
          sP[(tl * V + v) * PP + pi * a.p + pj] = px;
          sc = fmaf(px, ts[pi * a.p + pj], sc);
        }
    } else {
      for (int k = 0; k < PP; ++k) sP[(tl * V + v) * PP + k] = 0.f;
    }
    ssc[tl * V + v] = sc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TT * V; i += NT) {
    const int tl = i % TT, v = i / TT;
    float w = 0.f;
    if (t0 + tl < a.T) {
      float mx = -INFINITY;
      for (int u = 0; u < V; ++u) mx = fmaxf(mx, ssc[tl * V + u]);
      float sum = 0.f;
      for (int u = 0; u < V; ++u) sum += __expf(ssc[tl * V + u] - mx);
      w = __expf(ssc[tl * V + v] - mx) / sum;
    }
    sa[tl * V + v] = w;
  }
  __syncthreads();
}

template <typename T, int HD>
__global__ void __launch_bounds__(NT) frontend_fwd_kernel(const FeArgs a) {
  constexpr int EPT = HD / 4;
  extern __shared__ float smem[];
  const int V = a.V, PP = a.PP, P1 = PP + 1, KK = a.KK;
  float* sM = smem;                      // [KK][HD]
  float* sP = sM + KK * HD;              // [TT][V][PP]
  float* sa = sP + TT * V * PP;          // [TT][V]
  float* ssc = sa + TT * V;              // [TT][V]
  const int h = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * TT;
  for (int i = threadIdx.x; i < KK * HD; i += NT) sM[i] = a.tab_v[(size_t)h * KK * HD + i];
  stage_tile(a, h, t0, sP, sa, ssc);
  const int tl = threadIdx.x >> 2, eg = threadIdx.x & 3;
  float acc[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) acc[e] = 0.f;
  for (int v = 0; v < V; ++v) {
    const float w = sa[tl * V + v];
    for (int k = 0; k < P1; ++k) {
      const float c = (k < PP) ? w * sP[(tl * V + v) * PP + k] : w;
      const float* mrow = sM + (v * P1 + k) * HD + eg * EPT;
#pragma unroll
      for (int e = 0; e < EPT; e += 4) {
        const float4 m4 = *reinterpret_cast<const float4*>(mrow + e);
        acc[e] = fmaf(c, m4.x, acc[e]); acc[e + 1] = fmaf(c, m4.y, acc[e + 1]);
        acc[e + 2] = fmaf(c, m4.z, acc[e + 2]); acc[e + 3] = fmaf(c, m4.w, acc[e + 3]);
      }
    }
  }
  const long long t = t0 + tl;
  if (t < a.T) {
    T* op = reinterpret_cast<T*>(a.out) + (size_t)t * a.heads * HD + (size_t)h * HD + eg * EPT;
#pragma unroll
    for (int e = 0; e < EPT; ++e) op[e] = from_f<T>(acc[e]);
  }
}

// Backward of one head over persistent token tiles.  Per tile (TT tokens) two register-tiled fp32 contractions
//   d(tab_v)[kk][e] += sum_t C[t][kk] dO[t][e]      C[t][kk] = a[t][v] P'[t][v][k']   (kk = v*(PP+1)+k')
//   dC[t][kk]        = sum_e dO[t][e] tab_v[kk][e]
// then da = <dC, P'>, ds = a (da - <a, da>) and d(tab_s)[v][k'] += sum_t ds[t][v] P'[t][v][k'].
constexpr int KKP = 128;     // padded table rows (KK <= 128)
constexpr int DOP = 4;       // padding of the dO rows (bank spread for the dC contraction)

template <typename T, int HD>
__global__ void __launch_bounds__(NT) frontend_bwd_kernel(const FeArgs a) {
  extern __shared__ float smem[];
  const int V = a.V, PP = a.PP, P1 = PP + 1, KK = a.KK;
  float* sM = smem;                      // [KKP][HD]   (rows >= KK zero)
  float* sP = sM + KKP * HD;             // [TT][V][PP]
  float* sa = sP + TT * V * PP;          // [TT][V]
  float* ssc = sa + TT * V;              // [TT][V]  (scores, then da, then ds)
  float* sdO = ssc + TT * V;             // [TT][HD + DOP]
  float* sC = sdO + TT * (HD + DOP);     // [TT][KKP]  C, then dC
  const int h = blockIdx.y;
  for (int i = threadIdx.x; i < KKP * HD; i += NT) sM[i] = (i < KK * HD) ? a.tab_v[(size_t)h * KK * HD + i] : 0.f;

  // d(tab_v) slice of this thread: rows kt*KROWS .. +KROWS, columns et*4 .. +3
  constexpr int ETH = HD / 4;
  constexpr int KTH = NT / ETH;
  constexpr int KROWS = KKP / KTH;
  const int et = threadIdx.x % ETH, kt = threadIdx.x / ETH;
  float dM[KROWS][4];
#pragma unroll
  for (int r = 0; r < KROWS; ++r) { dM[r][0] = dM[r][1] = dM[r][2] = dM[r][3] = 0.f; }
  float dS = 0.f;                         // d(tab_s)[v,h,k'] for kk = threadIdx.x (< KK)
  // dC tile of this thread: tokens tg*4 .. +3, rows kg*8 .. +7
  const int tg = threadIdx.x % (TT / 4), kg = threadIdx.x / (TT / 4);

  const long long ntiles = (a.T + TT - 1) / TT;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long t0 = tile * TT;
    __syncthreads();
    stage_tile(a, h, t0, sP, sa, ssc);
    for (int i = threadIdx.x; i < TT * (HD / 4); i += NT) {
      const int tl = i / (HD / 4), e4 = i % (HD / 4);
      const long long t = t0 + tl;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < a.T) {
        const T* src = reinterpret_cast<const T*>(a.dout) + (size_t)t * a.heads * HD + (size_t)h * HD + e4 * 4;
        v = make_float4(to_f(src[0]), to_f(src[1]), to_f(src[2]), to_f(src[3]));
      }
      *reinterpret_cast<float4*>(sdO + tl * (HD + DOP) + e4 * 4) = v;
    }
    for (int i = threadIdx.x; i < TT * KKP; i += NT) {
      const int tl = i / KKP, kk = i % KKP;
      float c = 0.f;
      if (kk < KK) {
        const int v = kk / P1, k = kk % P1;
        const float w = sa[tl * V + v];
        c = (k < PP) ? w * sP[(tl * V + v) * PP + k] : w;
      }
      sC[i] = c;
    }
    __syncthreads();
    // d(tab_v) += C^T dO
#pragma unroll 2
    for (int tl = 0; tl < TT; ++tl) {
      const float4 d4 = *reinterpret_cast<const float4*>(sdO + tl * (HD + DOP) + et * 4);
      float c[KROWS];
#pragma unroll
      for (int r = 0; r < KROWS; r += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(sC + tl * KKP + kt * KROWS + r);
        c[r] = c4.x; c[r + 1] = c4.y; c[r + 2] = c4.z; c[r + 3] = c4.w;
      }
#pragma unroll
      for (int r = 0; r < KROWS; ++r) {
        dM[r][0] = fmaf(c[r], d4.x, dM[r][0]); dM[r][1] = fmaf(c[r], d4.y, dM[r][1]);
        dM[r][2] = fmaf(c[r], d4.z, dM[r][2]); dM[r][3] = fmaf(c[r], d4.w, dM[r][3]);
      }
    }
    // dC = dO tab_v^T (4 tokens x 8 rows per thread)
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int e = 0; e < HD; e += 4) {
      float4 d[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = *reinterpret_cast<const float4*>(sdO + (tg * 4 + i) * (HD + DOP) + e);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 m4 = *reinterpret_cast<const float4*>(sM + (kg * 8 + j) * HD + e);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          acc[i][j] += d[i].x * m4.x + d[i].y * m4.y + d[i].z * m4.z + d[i].w * m4.w;
      }
    }
    __syncthreads();                       // everyone is done reading C
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float4*>(sC + (tg * 4 + i) * KKP + kg * 8) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(sC + (tg * 4 + i) * KKP + kg * 8 + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    __syncthreads();
    // da[t][v] = sum_k' dC[t][v,k'] P'[t][v][k']
    for (int i = threadIdx.x; i < TT * V; i += NT) {
      const int tl = i % TT, v = i / TT;
      float da = sC[tl * KKP + v * P1 + PP];
      for (int k = 0; k < PP; ++k) da = fmaf(sC[tl * KKP + v * P1 + k], sP[(tl * V + v) * PP + k], da);
      ssc[tl * V + v] = da;
    }
    __syncthreads();
    float myds[(TT * 32 + NT - 1) / NT];   // V <= 32
    {
      int n = 0;
      for (int i = threadIdx.x; i < TT * V; i += NT, ++n) {
        const int tl = i % TT, v = i / TT;
        float dot = 0.f;
        for (int u = 0; u < V; ++u) dot = fmaf(sa[tl * V + u], ssc[tl * V + u], dot);
        myds[n] = sa[tl * V + v] * (ssc[tl * V + v] - dot);
      }
    }
    __syncthreads();
    {
      int n = 0;
      for (int i = threadIdx.x; i < TT * V; i += NT, ++n) ssc[i % TT * V + i / TT] = myds[n];
    }
    __syncthreads();
    if (threadIdx.x < KK) {
      const int v = threadIdx.x / P1, k = threadIdx.x % P1;
      float sum = 0.f;
      for (int tl = 0; tl < TT; ++tl) {
        const float pk = (k < PP) ? sP[(tl * V + v) * PP + k] : 1.f;
        sum = fmaf(ssc[tl * V + v], pk, sum);
      }
      dS += sum;
    }
  }
  if (threadIdx.x < KK) {
    const int v = threadIdx.x / P1, k = threadIdx.x % P1;
    atomicAdd(&a.dtab_s[((size_t)v * a.heads + h) * P1 + k], dS);
  }
#pragma unroll
  for (int r = 0; r < KROWS; ++r) {
    const int kk = kt * KROWS + r;
    if (kk < KK) {
      float* dst = a.dtab_v + ((size_t)h * KK + kk) * HD + et * 4;
      atomicAdd(dst + 0, dM[r][0]); atomicAdd(dst + 1, dM[r][1]);
      atomicAdd(dst + 2, dM[r][2]); atomicAdd(dst + 3, dM[r][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ bf16 arm: HMMA
// The two big contractions of the collapsed front end are small-K GEMMs against the per-head table (K = V*(PP+1) <= 128):
//   forward   o[t, e]   = sum_kk C[t, kk] tab_v[kk, e]            backward  dC[t, kk] = sum_e dO[t, e] tab_v[kk, e]
//                                                                            dM[kk, e] += sum_t C[t, kk] dO[t, e]
// In the bf16 arm they run on the tensor cores with warp-level mma.sync (m16n8k16, bf16 in, fp32 accumulate): the work is
// ~30 GFLOP per pass -- far too little to justify a tcgen05/TMEM pipeline, and at mma.sync speed the kernels are bound by
// the x read and the o / dO traffic.  The fp32 arm keeps the exact SIMT kernels above.
constexpr int LDS_ = 136;          // bf16 row pitch of the smem operand tiles (128 + 8: conflict-free fragment loads)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a_)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a_[0]), "r"(a_[1]), "r"(a_[2]), "r"(a_[3]), "r"(b0), "r"(b1));
}

// pixels of one token tile for all variables (head independent)
__device__ __forceinline__ void stage_pixels(const FeArgs& a, long long t0, float* sP) {
  const int V = a.V, PP = a.PP, L = a.gh * a.gw;
  if (a.p == 2) {
    // p = 2: a thread keeps its token (NT is a multiple of TT) and walks the variables: one coordinate computation, two
    // 8-byte loads per (token, variable) when the rows are 8-byte aligned, one 16-byte shared store
    static_assert(NT % TT == 0, "a thread must keep its token across the variable loop");
    const int tl = threadIdx.x % TT;
    const long long t = t0 + tl;
    float4* dst = reinterpret_cast<float4*>(sP) + tl * V;
    if (t < a.T) {
      const int b = (int)(t / L), l = (int)(t - (long long)b * L);
      const int gy = l / a.gw, gx = l - gy * a.gw;
      const size_t plane = (size_t)a.Hx * a.Wx;
      const float* xp = a.x + (size_t)b * V * plane + (size_t)(2 * gy) * a.Wx + 2 * gx;
      const bool vec = ((a.Wx & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 7) == 0);
      for (int v = threadIdx.x / TT; v < V; v += NT / TT) {
        const float* q = xp + (size_t)v * plane;
        float4 px;
        if (vec) {
          const float2 r0 = __ldg(reinterpret_cast<const float2*>(q)), r1 = __ldg(reinterpret_cast<const float2*>(q + a.Wx));
          px = make_float4(r0.x, r0.y, r1.x, r1.y);
        } else {
          px = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + a.Wx), __ldg(q + a.Wx + 1));
        }
        dst[v] = px;
      }
    } else {
      for (int v = threadIdx.x / TT; v < V; v += NT / TT) dst[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  for (int i = threadIdx.x; i < TT * V; i += NT) {
    const int tl = i % TT, v = i / TT;
    const long long t = t0 + tl;
    if (t < a.T) {
      const int b = (int)(t / L), l = (int)(t % L);
      const int gy = l / a.gw, gx = l % a.gw;
      const float* xp = a.x + (((size_t)b * V + v) * a.Hx + (size_t)gy * a.p) * a.Wx + (size_t)gx * a.p;
      for (int pi = 0; pi < a.p; ++pi)
        for (int pj = 0; pj < a.p; ++pj) sP[(tl * V + v) * PP + pi * a.p + pj] = xp[(size_t)pi * a.Wx + pj];
    } else {
      for (int k = 0; k < PP; ++k) sP[(tl * V + v) * PP + k] = 0.f;
    }
  }
}

constexpr int LDB_ = 72;           // bf16 row pitch of the [kk][e] table tile (64 + 8)

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(smem_row)));
}

// softmax weights a[t][v] of head h from the staged pixels (NT == 4 * TT: four threads per token), then the bf16 coefficient tile
// sC[t][kk] = a[t][v] * [pixels, 1] (kk = v*(PP+1)+k'; columns >= KK stay zero from the caller's one-time clear)
// this head's score table tab_s[:, h, :] -> shared memory [V][5] (fast path only: V * 5 <= 128 floats)
__device__ __forceinline__ void stage_tab_s(const FeArgs& a, int h, float* sTs) {
  for (int i = threadIdx.x; i < a.V * 5; i += NT) sTs[i] = __ldg(a.tab_s + ((size_t)(i / 5) * a.heads + h) * 5 + i % 5);
}

// sTs holds head h's score table on entry (fast path); h_next >= 0: it is refilled for that head once the scores are done
__device__ __forceinline__ void head_coefficients(const FeArgs& a, int h, long long t0, const float* sP, float* sa,
                                                  __nv_bfloat16* sC, float* sTs, int h_next) {
  const int V = a.V, PP = a.PP, P1 = PP + 1;
  if (a.fast) {
    // p = 2 (PP = 4): pixels as one 16-byte load per (token, variable), coefficients of two neighbouring variables packed
    // into one bf16x2 store per k' (columns kk' = k' * VP + v, so that the pair is adjacent), everything unrolled
    const float4* sP4 = reinterpret_cast<const float4*>(sP);
    {
      const int tl = threadIdx.x >> 2, sub = threadIdx.x & 3;
      float mx = -INFINITY;
      for (int v = sub; v < V; v += 4) {
        const float* ts = sTs + v * 5;
        const float4 px = sP4[tl * V + v];
        float sc = ts[4];
        sc = fmaf(px.x, ts[0], sc);
        sc = fmaf(px.y, ts[1], sc);
        sc = fmaf(px.z, ts[2], sc);
        sc = fmaf(px.w, ts[3], sc);
        sa[tl * V + v] = sc;
        mx = fmaxf(mx, sc);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
      for (int v = sub; v < V; v += 4) {
        const float e = __expf(sa[tl * V + v] - mx);
        sa[tl * V + v] = e;
        sum += e;
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = (t0 + tl < a.T) ? 1.f / sum : 0.f;
      for (int v = sub; v < V; v += 4) sa[tl * V + v] *= inv;
    }
    __syncthreads();
    if (h_next >= 0) stage_tab_s(a, h_next, sTs);       // visible to the next head's scores through the barriers in between
    const int VP = a.VP, VH = VP >> 1;
    for (int i = threadIdx.x; i < TT * VH; i += NT) {
      const int vp = i % VH, tl = i / VH;
      const int v0 = 2 * vp, v1 = v0 + 1;
      const bool has1 = v1 < V;
      const float w0 = sa[tl * V + v0], w1 = has1 ? sa[tl * V + v1] : 0.f;
      const float4 p0 = sP4[tl * V + v0];
      const float4 p1 = has1 ? sP4[tl * V + v1] : make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t* c = reinterpret_cast<uint32_t*>(sC + tl * LDS_) + vp;
      c[0 * VH] = pack_bf16x2(w0 * p0.x, w1 * p1.x);
      c[1 * VH] = pack_bf16x2(w0 * p0.y, w1 * p1.y);
      c[2 * VH] = pack_bf16x2(w0 * p0.z, w1 * p1.z);
      c[3 * VH] = pack_bf16x2(w0 * p0.w, w1 * p1.w);
      c[4 * VH] = pack_bf16x2(w0, w1);
    }
    return;
  }
  {   // four threads per token: scores of every 4th variable, max / sum combined with two shuffles
    const int tl = threadIdx.x >> 2, sub = threadIdx.x & 3;
    float mx = -INFINITY;
    for (int v = sub; v < V; v += 4) {
      const float* ts = a.tab_s + ((size_t)v * a.heads + h) * P1;
      float sc = ts[PP];
      for (int k = 0; k < PP; ++k) sc = fmaf(sP[(tl * V + v) * PP + k], ts[k], sc);
      sa[tl * V + v] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
    for (int v = sub; v < V; v += 4) {
      const float e = __expf(sa[tl * V + v] - mx);
      sa[tl * V + v] = e;
      sum += e;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = (t0 + tl < a.T) ? 1.f / sum : 0.f;
    for (int v = sub; v < V; v += 4) sa[tl * V + v] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TT * V; i += NT) {
    const int tl = i % TT, v = i / TT;
    const float w = sa[tl * V + v];
    __nv_bfloat16* c = sC + tl * LDS_ + v * P1;
    for (int k = 0; k < PP; ++k) c[k] = __float2bfloat16_rn(w * sP[(tl * V + v) * PP + k]);
    c[PP] = __float2bfloat16_rn(w);
  }
}

// o[t, h*64 + e] for a tile of 64 tokens and ALL heads (the pixels are staged once per tile)
__global__ void __launch_bounds__(NT) frontend_fwd_mma_kernel(const FeArgs a) {
  constexpr int HD = 64;
  extern __shared__ float smem[];
  const int V = a.V, PP = a.PP, KK = a.KK;
  float* sP = smem;                                   // [TT][V][PP]
  float* sa = sP + TT * V * PP;                       // [TT][V]
  __nv_bfloat16* sC = reinterpret_cast<__nv_bfloat16*>(sa + TT * V);   // [TT][LDS_]  coefficients
  __nv_bfloat16* sB = sC + TT * LDS_;                 // [KKP][LDB_] tab_v[h] (row kk, e contiguous; rows >= KK zero)
  __nv_bfloat16* sO = sB + KKP * LDB_;                // [TT][LDB_]  output tile
  float* sTs = reinterpret_cast<float*>(sO + TT * LDB_);   // [V][5] score table of the current head (fast path)
  const long long t0 = (long long)blockIdx.x * TT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  const int mt = warp & 3, nh = warp >> 2;            // 16-token row tile, 32-column half
  stage_pixels(a, t0, sP);
  if (a.fast) stage_tab_s(a, 0, sTs);
  for (int i = threadIdx.x; i < TT * LDS_ / 2; i += NT) reinterpret_cast<uint32_t*>(sC)[i] = 0u;
  for (int i = threadIdx.x; i < KKP * LDB_ / 2; i += NT) reinterpret_cast<uint32_t*>(sB)[i] = 0u;
  __syncthreads();
  for (int h = 0; h < a.heads; ++h) {
    head_coefficients(a, h, t0, sP, sa, sC, sTs, h + 1 < a.heads ? h + 1 : -1);
    for (int i = threadIdx.x; i < KK * (HD / 4); i += NT) {          // fp32 table -> bf16 tile, 16-byte reads
      const int kk = i / (HD / 4), e4 = i % (HD / 4);
      const float4 v4 = *reinterpret_cast<const float4*>(a.tab_v + ((size_t)h * KK + kk) * HD + e4 * 4);
      uint2 pk;
      pk.x = pack_bf16x2(v4.x, v4.y);
      pk.y = pack_bf16x2(v4.z, v4.w);
      *reinterpret_cast<uint2*>(sB + (a.fast ? (kk % 5) * a.VP + kk / 5 : kk) * LDB_ + e4 * 4) = pk;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KKP / 16; ++ks) {
      const int k0 = ks * 16 + tig * 2;
      uint32_t af[4];
      af[0] = *reinterpret_cast<const uint32_t*>(sC + (mt * 16 + g) * LDS_ + k0);
      af[1] = *reinterpret_cast<const uint32_t*>(sC + (mt * 16 + g + 8) * LDS_ + k0);
      af[2] = *reinterpret_cast<const uint32_t*>(sC + (mt * 16 + g) * LDS_ + k0 + 8);
      af[3] = *reinterpret_cast<const uint32_t*>(sC + (mt * 16 + g + 8) * LDS_ + k0 + 8);
      // B fragments of two n-tiles per ldmatrix.x4.trans: matrices (k 0-7 | 8-15) x (n 0-7 | 8-15) of the row-major table
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, sB + (ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDB_ + nh * 32 + np * 16 + (lane >> 4) * 8);
        mma_bf16_16816(acc[np * 2], af, bf[0], bf[1]);
        mma_bf16_16816(acc[np * 2 + 1], af, bf[2], bf[3]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = nh * 32 + nt * 8 + tig * 2;
      *reinterpret_cast<uint32_t*>(sO + (mt * 16 + g) * LDB_ + col) = pack_bf16x2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<uint32_t*>(sO + (mt * 16 + g + 8) * LDB_ + col) = pack_bf16x2(acc[nt][2], acc[nt][3]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TT * (HD / 8); i += NT) {      // 128-byte rows, 16 bytes per thread
      const int tl = i / (HD / 8), c8 = i % (HD / 8);
      if (t0 + tl < a.T)
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + (size_t)(t0 + tl) * a.heads * HD + (size_t)h * HD + c8 * 8) =
            *reinterpret_cast<const uint4*>(sO + tl * LDB_ + c8 * 8);
    }
    // (the next head's coefficient / table writes are ordered behind this head's reads by the barrier inside
    //  head_coefficients and the one after the table copy; sO is rewritten only after the next MMA phase)
  }
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(smem_row)));
}

// Backward of one head over persistent token tiles, bf16 arm.  Per tile two tensor-core contractions
//   dM[kk][e] += sum_t C[t][kk] dO[t][e]   (warp w owns rows 16w..16w+15 of dM in registers for the whole kernel)
//   dC[t][kk]  = sum_e dO[t][e] M[kk][e]
// then the small SIMT tail (da, ds, d tab_s) of the fp32 kernel.
__global__ void __launch_bounds__(NT, 2) frontend_bwd_mma_kernel(const FeArgs a) {
  constexpr int HD = 64;
  constexpr int LDC_ = KKP + 4;                       // fp32 pitch of the dC tile
  extern __shared__ float smem[];
  const int V = a.V, PP = a.PP, P1 = PP + 1, KK = a.KK;
  float* sP = smem;                                   // [TT][V][PP]
  float* sa = sP + TT * V * PP;                       // [TT][V]
  float* ssc = sa + TT * V;                           // [TT][V]
  float* sdC = ssc + TT * V;                          // [TT][LDC_] fp32; its head doubles as the bf16 coefficient tile sC
  __nv_bfloat16* sC = reinterpret_cast<__nv_bfloat16*>(sdC);
  __nv_bfloat16* sdO = reinterpret_cast<__nv_bfloat16*>(sdC + TT * LDC_);   // [TT][LDB_]
  __nv_bfloat16* sM = sdO + TT * LDB_;                // [KKP][LDB_]  tab_v[h] (rows >= KK zero)
  float* sTs = reinterpret_cast<float*>(sM + KKP * LDB_);   // [V][5] score table of this head (fast path)
  const int h = blockIdx.y;
  if (a.fast) stage_tab_s(a, h, sTs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  for (int i = threadIdx.x; i < KKP * LDB_ / 2; i += NT) reinterpret_cast<uint32_t*>(sM)[i] = 0u;
  __syncthreads();
  for (int i = threadIdx.x; i < KK * (HD / 4); i += NT) {
    const int kk = i / (HD / 4), e4 = i % (HD / 4);
    const float4 v4 = *reinterpret_cast<const float4*>(a.tab_v + ((size_t)h * KK + kk) * HD + e4 * 4);
    uint2 pk;
    pk.x = pack_bf16x2(v4.x, v4.y);
    pk.y = pack_bf16x2(v4.z, v4.w);
    *reinterpret_cast<uint2*>(sM + (a.fast ? (kk % 5) * a.VP + kk / 5 : kk) * LDB_ + e4 * 4) = pk;
  }
  float dM[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { dM[nt][0] = dM[nt][1] = dM[nt][2] = dM[nt][3] = 0.f; }
  float dS = 0.f;
  const int mt = warp & 3, nq = warp >> 2;            // dC: 16-token row tile, 64-column half of kk

  const long long ntiles = (a.T + TT - 1) / TT;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long t0 = tile * TT;
    __syncthreads();                                  // previous tile's tail is done with sP / sa / sdC
    stage_pixels(a, t0, sP);
    for (int i = threadIdx.x; i < TT * LDS_ / 2; i += NT) reinterpret_cast<uint32_t*>(sC)[i] = 0u;
    for (int i = threadIdx.x; i < TT * (HD / 8); i += NT) {
      const int tl = i / (HD / 8), c8 = i % (HD / 8);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (t0 + tl < a.T)
        v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.dout) + (size_t)(t0 + tl) * a.heads * HD +
                                            (size_t)h * HD + c8 * 8);
      *reinterpret_cast<uint4*>(sdO + tl * LDB_ + c8 * 8) = v;
    }
    __syncthreads();
    head_coefficients(a, h, t0, sP, sa, sC, sTs, -1);
    __syncthreads();
    // ---- dM += C^T dO   (M = kk rows 16*warp.., N = e, K = t)
#pragma unroll
    for (int ks = 0; ks < TT / 16; ++ks) {
      uint32_t af[4];   // A = C^T: matrices (kk 0-7 | 8-15) x (t 0-7 | 8-15), transposed out of the row-major [t][kk] tile
      ldmatrix_x4_trans(af, sC + (ks * 16 + (lane >> 4) * 8 + (lane & 7)) * LDS_ + warp * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, sdO + (ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDB_ + np * 16 + (lane >> 4) * 8);
        mma_bf16_16816(dM[np * 2], af, bf[0], bf[1]);
        mma_bf16_16816(dM[np * 2 + 1], af, bf[2], bf[3]);
      }
    }
    // ---- dC = dO M^T   (M = t rows 16*mt.., N = kk in [64 nq, +64), K = e)
    float dc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { dc[nt][0] = dc[nt][1] = dc[nt][2] = dc[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      uint32_t af[4];   // A = dO row-major: matrices (t 0-7 | 8-15) x (e 0-7 | 8-15)
      ldmatrix_x4(af, sdO + (mt * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDB_ + ks * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const __nv_bfloat16* bp = sM + (nq * 64 + nt * 8 + g) * LDB_ + ks * 16 + tig * 2;
        mma_bf16_16816(dc[nt], af, *reinterpret_cast<const uint32_t*>(bp), *reinterpret_cast<const uint32_t*>(bp + 8));
      }
    }
    __syncthreads();                                  // all reads of sC are done: overwrite it with the fp32 dC tile
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nq * 64 + nt * 8 + tig * 2;
      *reinterpret_cast<float2*>(sdC + (mt * 16 + g) * LDC_ + col) = make_float2(dc[nt][0], dc[nt][1]);
      *reinterpret_cast<float2*>(sdC + (mt * 16 + g + 8) * LDC_ + col) = make_float2(dc[nt][2], dc[nt][3]);
    }
    __syncthreads();
    // ---- SIMT tail: da[t][v] = <dC[t][v,:], P'[t][v,:]>,  ds = a (da - <a, da>),  d tab_s[v][k'] += sum_t ds P'
    for (int i = threadIdx.x; i < TT * V; i += NT) {
      const int tl = i % TT, v = i / TT;
      float da;
      if (a.fast) {
        const float* dc_ = sdC + tl * LDC_ + v;
        const float4 px = reinterpret_cast<const float4*>(sP)[tl * V + v];
        da = dc_[4 * a.VP];
        da = fmaf(dc_[0], px.x, da);
        da = fmaf(dc_[a.VP], px.y, da);
        da = fmaf(dc_[2 * a.VP], px.z, da);
        da = fmaf(dc_[3 * a.VP], px.w, da);
      } else {
        da = sdC[tl * LDC_ + v * P1 + PP];
        for (int k = 0; k < PP; ++k) da = fmaf(sdC[tl * LDC_ + v * P1 + k], sP[(tl * V + v) * PP + k], da);
      }
      ssc[tl * V + v] = da;
    }
    __syncthreads();
    {
      const int tl = threadIdx.x >> 2, sub = threadIdx.x & 3;     // four threads per token
      float dot = 0.f;
      for (int v = sub; v < V; v += 4) dot = fmaf(sa[tl * V + v], ssc[tl * V + v], dot);
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      for (int v = sub; v < V; v += 4) ssc[tl * V + v] = sa[tl * V + v] * (ssc[tl * V + v] - dot);
    }
    __syncthreads();
    if ((threadIdx.x & 127) < KK) {                   // two threads per table entry, half of the tile's tokens each
      const int vk = threadIdx.x & 127, v = vk / P1, k = vk % P1;
      const int tl0 = (threadIdx.x >> 7) * (TT / 2);
      float sum = 0.f;
#pragma unroll 4
      for (int tl = tl0; tl < tl0 + TT / 2; ++tl) {
        const float pk = (k < PP) ? sP[(tl * V + v) * PP + k] : 1.f;
        sum = fmaf(ssc[tl * V + v], pk, sum);
      }
      dS += sum;
    }
  }
  if ((threadIdx.x & 127) < KK) {
    const int vk = threadIdx.x & 127, v = vk / P1, k = vk % P1;
    atomicAdd(&a.dtab_s[((size_t)v * a.heads + h) * P1 + k], dS);
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + tig * 2;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      int kk = warp * 16 + g + half * 8;
      if (a.fast) {                                   // row kk' = k' * VP + v of the permuted tile -> table row v * 5 + k'
        const int kq = kk / a.VP, v = kk % a.VP;
        kk = (kq < 5 && v < V) ? v * 5 + kq : KK;
      }
      if (kk < KK) {
        float* dst = a.dtab_v + ((size_t)h * KK + kk) * HD + col;
        atomicAdd(dst, dM[nt][half * 2]);
        atomicAdd(dst + 1, dM[nt][half * 2 + 1]);
      }
    }
  }
}

int fill(FeArgs& a, const float* x, const float* tab_s, const float* tab_v, int B, int V, int Hx, int Wx, int p, int gh,
         int gw, int heads, int hd) {
  O2_REQUIRE(x && tab_s && tab_v, "frontend: null pointer");
  O2_REQUIRE(B > 0 && V > 0 && V <= 32 && p > 0 && gh > 0 && gw > 0 && heads > 0, "frontend: bad dims");
  O2_REQUIRE(gh * p <= Hx && gw * p <= Wx, "frontend: token grid %dx%d (p=%d) exceeds the field %dx%d", gh, gw, p, Hx, Wx);
  O2_REQUIRE(hd == 32 || hd == 64 || hd == 128, "frontend: head dim %d not in {32,64,128}", hd);
  O2_REQUIRE(V * (p * p + 1) <= 128, "frontend: V*(p*p+1)=%d exceeds 128", V * (p * p + 1));
  O2_REQUIRE(heads <= 65535, "frontend: too many heads");
  memset(&a, 0, sizeof(a));
  a.x = x; a.tab_s = tab_s; a.tab_v = tab_v;
  a.B = B; a.V = V; a.Hx = Hx; a.Wx = Wx; a.p = p; a.gh = gh; a.gw = gw; a.heads = heads; a.hd = hd;
  a.PP = p * p; a.KK = V * (p * p + 1);
  a.VP = (V + 1) & ~1;
  a.fast = (a.PP == 4 && 5 * a.VP <= 128) ? 1 : 0;   // read by the bf16 (mma) kernels only
  a.T = (long long)B * gh * gw;
  return O2_OK;
}

template <typename T, int HD> int launch_fwd(const FeArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)a.KK * HD + (size_t)TT * a.V * a.PP + 2 * (size_t)TT * a.V);
  O2_CUDA(cudaFuncSetAttribute(frontend_fwd_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((a.T + TT - 1) / TT), a.heads);
  frontend_fwd_kernel<T, HD><<<grid, NT, smem, st>>>(a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}
template <typename T, int HD> int launch_bwd(const FeArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)KKP * HD + (size_t)TT * a.V * a.PP + 2 * (size_t)TT * a.V +
                                       (size_t)TT * (HD + DOP) + (size_t)TT * KKP);
  O2_REQUIRE(smem <= 227 * 1024, "frontend_bwd: %zu bytes of shared memory needed", smem);
  O2_CUDA(cudaFuncSetAttribute(frontend_bwd_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long gx = (long long)o2_num_sms() / a.heads;     // one CTA per SM (shared memory), a single wave
  if (gx < 1) gx = 1;
  const long long ntiles = (a.T + TT - 1) / TT;
  if (gx > ntiles) gx = ntiles;
  frontend_bwd_kernel<T, HD><<<dim3((unsigned)gx, a.heads), NT, smem, st>>>(a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

}  // namespace

#define O2_FE_DISPATCH(FN, a, dtype, st)                                                     \
  do {                                                                                       \
    if ((dtype) == O2_F32) {                                                                 \
      if ((a).hd == 32) return FN<float, 32>(a, st);                                         \
      if ((a).hd == 64) return FN<float, 64>(a, st);                                         \
      return FN<float, 128>(a, st);                                                          \
    } else if ((dtype) == O2_BF16) {                                                         \
      if ((a).hd == 32) return FN<__nv_bfloat16, 32>(a, st);                                 \
      if ((a).hd == 64) return FN<__nv_bfloat16, 64>(a, st);                                 \
      return FN<__nv_bfloat16, 128>(a, st);                                                  \
    }                                                                                        \
    O2_FAIL(O2_ERR_ARG, "frontend: bad dtype %d", (dtype));                                  \
  } while (0)

extern "C" int o2_frontend_fwd(const float* x, const float* tab_s, const float* tab_v, void* out, int out_dtype, int B,
                               int V, int Hx, int Wx, int p, int gh, int gw, int heads, int hd, void* stream) {
  FeArgs a;
  int rc = fill(a, x, tab_s, tab_v, B, V, Hx, Wx, p, gh, gw, heads, hd);
  if (rc) return rc;
  O2_REQUIRE(out, "frontend_fwd: null out");
  a.out = out;
  if (out_dtype == O2_BF16 && hd == 64 && ((uintptr_t)out % 16) == 0 && !getenv("O2_FRONTEND_SIMT")) {
    const size_t smem = sizeof(float) * ((size_t)TT * a.V * a.PP + (size_t)TT * a.V + 128) + 2 * ((size_t)TT * LDS_ + (size_t)(KKP + TT) * LDB_);
    O2_CUDA(cudaFuncSetAttribute(frontend_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    frontend_fwd_mma_kernel<<<(unsigned)((a.T + TT - 1) / TT), NT, smem, (cudaStream_t)stream>>>(a);
    O2_LAUNCH_CHECK();
    return O2_OK;
  }
  O2_FE_DISPATCH(launch_fwd, a, out_dtype, (cudaStream_t)stream);
}

extern "C" int o2_frontend_bwd(const float* x, const float* tab_s, const float* tab_v, const void* dout, int dtype,
                               float* dtab_s, float* dtab_v, int B, int V, int Hx, int Wx, int p, int gh, int gw,
                               int heads, int hd, void* stream) {
  FeArgs a;
  int rc = fill(a, x, tab_s, tab_v, B, V, Hx, Wx, p, gh, gw, heads, hd);
  if (rc) return rc;
  O2_REQUIRE(dout && dtab_s && dtab_v, "frontend_bwd: null pointer");
  a.dout = dout; a.dtab_s = dtab_s; a.dtab_v = dtab_v;
  if (dtype == O2_BF16 && hd == 64 && ((uintptr_t)dout % 16) == 0 && !getenv("O2_FRONTEND_SIMT")) {
    const size_t smem = sizeof(float) * ((size_t)TT * a.V * a.PP + 2 * (size_t)TT * a.V + (size_t)TT * (KKP + 4) + 128) +
                        2 * ((size_t)TT + KKP) * LDB_;
    O2_CUDA(cudaFuncSetAttribute(frontend_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (a.T + TT - 1) / TT;
    long long gx = (long long)o2_num_sms() * 2 / a.heads;
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    frontend_bwd_mma_kernel<<<dim3((unsigned)gx, a.heads), NT, smem, (cudaStream_t)stream>>>(a);
    O2_LAUNCH_CHECK();
    return O2_OK;
  }
  O2_FE_DISPATCH(launch_bwd, a, dtype, (cudaStream_t)stream);
}
