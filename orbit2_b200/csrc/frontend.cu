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
#include "common.cuh"

namespace {

constexpr int TT = 64;      // tokens per tile
constexpr int NT = 256;
constexpr int KPT = 8;      // table rows per thread in the d(tab_v) accumulation -> KK <= 128

struct FeArgs {
  const float* x; const float* tab_s; const float* tab_v;
  void* out; const void* dout; float* dtab_s; float* dtab_v;
  int B, V, Hx, Wx, p, gh, gw, heads, hd, PP, KK;
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

template <typename T, int HD>
__global__ void __launch_bounds__(NT) frontend_bwd_kernel(const FeArgs a) {
  extern __shared__ float smem[];
  const int V = a.V, PP = a.PP, P1 = PP + 1, KK = a.KK;
  float* sM = smem;                      // [KK][HD]
  float* sP = sM + KK * HD;              // [TT][V][PP]
  float* sa = sP + TT * V * PP;          // [TT][V]
  float* ssc = sa + TT * V;              // [TT][V]  (scores, then ds)
  float* sdO = ssc + TT * V;             // [TT][HD]
  float* sdC = sdO + TT * HD;            // [TT][KK]
  const int h = blockIdx.y;
  for (int i = threadIdx.x; i < KK * HD; i += NT) sM[i] = a.tab_v[(size_t)h * KK * HD + i];

  // d(tab_v) slice owned by this thread: rows kk = kt*KPT + r, columns e = et*4 .. +3 (for HD=64: 16 x 16 threads)
  constexpr int ETH = HD / 4;            // threads along e
  constexpr int KTH = NT / ETH;          // threads along kk
  const int et = threadIdx.x % ETH, kt = threadIdx.x / ETH;
  constexpr int KROWS = (128 + KTH - 1) / KTH;   // rows per thread to cover KK <= 128
  float dM[KROWS][4];
#pragma unroll
  for (int r = 0; r < KROWS; ++r) { dM[r][0] = dM[r][1] = dM[r][2] = dM[r][3] = 0.f; }
  float dS = 0.f;                         // d(tab_s)[v,h,k'] for kk = threadIdx.x (< KK)

  const long long ntiles = (a.T + TT - 1) / TT;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long t0 = tile * TT;
    __syncthreads();
    stage_tile(a, h, t0, sP, sa, ssc);
    for (int i = threadIdx.x; i < TT * HD; i += NT) {
      const int tl = i / HD, e = i % HD;
      const long long t = t0 + tl;
      sdO[i] = (t < a.T) ? to_f(reinterpret_cast<const T*>(a.dout)[(size_t)t * a.heads * HD + (size_t)h * HD + e]) : 0.f;
    }
    __syncthreads();
    // dC[t][kk] = dO[t][:] . tab_v[kk][:]
    for (int i = threadIdx.x; i < TT * KK; i += NT) {
      const int tl = i / KK, kk = i % KK;
      const float* d = sdO + tl * HD;
      const float* m = sM + kk * HD;
      float s = 0.f;
#pragma unroll 4
      for (int e = 0; e < HD; e += 4) {
        const float4 d4 = *reinterpret_cast<const float4*>(d + e);
        const float4 m4 = *reinterpret_cast<const float4*>(m + e);
        s += d4.x * m4.x + d4.y * m4.y + d4.z * m4.z + d4.w * m4.w;
      }
      sdC[tl * KK + kk] = s;
    }
    __syncthreads();
    // da[t][v] = sum_k' dC[t][v,k'] P'[t][v][k'];  ds = a (da - sum_u a_u da_u)
    for (int i = threadIdx.x; i < TT * V; i += NT) {
      const int tl = i % TT, v = i / TT;
      float da = sdC[tl * KK + v * P1 + PP];
      for (int k = 0; k < PP; ++k) da = fmaf(sdC[tl * KK + v * P1 + k], sP[(tl * V + v) * PP + k], da);
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
    // d(tab_s): one (v,k') per thread
    if (threadIdx.x < KK) {
      const int v = threadIdx.x / P1, k = threadIdx.x % P1;
      float s = 0.f;
      for (int tl = 0; tl < TT; ++tl) {
        const float pk = (k < PP) ? sP[(tl * V + v) * PP + k] : 1.f;
        s = fmaf(ssc[tl * V + v], pk, s);
      }
      dS += s;
    }
    // d(tab_v)[kk][e] += sum_t a[t][v] P'[t][v][k'] dO[t][e]
    for (int tl = 0; tl < TT; ++tl) {
      const float4 d4 = *reinterpret_cast<const float4*>(sdO + tl * HD + et * 4);
#pragma unroll
      for (int r = 0; r < KROWS; ++r) {
        const int kk = kt * KROWS + r;
        if (kk < KK) {
          const int v = kk / P1, k = kk % P1;
          const float w = sa[tl * V + v];
          const float c = (k < PP) ? w * sP[(tl * V + v) * PP + k] : w;
          dM[r][0] = fmaf(c, d4.x, dM[r][0]); dM[r][1] = fmaf(c, d4.y, dM[r][1]);
          dM[r][2] = fmaf(c, d4.z, dM[r][2]); dM[r][3] = fmaf(c, d4.w, dM[r][3]);
        }
      }
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
  const size_t smem = sizeof(float) * ((size_t)a.KK * HD + (size_t)TT * a.V * a.PP + 2 * (size_t)TT * a.V +
                                       (size_t)TT * HD + (size_t)TT * a.KK);
  O2_CUDA(cudaFuncSetAttribute(frontend_bwd_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long gx = ((long long)o2_num_sms() * 2 + a.heads - 1) / a.heads;
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
  O2_FE_DISPATCH(launch_bwd, a, dtype, (cudaStream_t)stream);
}
