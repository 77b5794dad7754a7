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
