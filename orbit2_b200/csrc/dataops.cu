// Either side of the hot path: the input normalisation of the data pipeline and the evaluation
// statistics, both single-pass HBM-bound kernels so that the host->device copy carries RAW fields and evaluation reads
// prediction and target once.
//
// o2_normalize_fields   reference: data/itermodule.py:202-211 (torchvision Normalize(mean, std) per variable, LogTransform
//                       for precipitation) applied per sample on the host by IndividualDataIter (iterdataset.py:360-379);
//                       LogTransform = precipmodule.py:21-42 (m -> mm, values <= 0.25 mm/day -> 0, log1p).
// o2_eval_stats         reference: metrics/functional.py:236-257 (rmse), :294-309 (pearson), :312-324 (mean_bias) and the
//                       denormalising TransformedMetric (metrics/metrics.py:100-115): per (sample, channel) fp64 sums of
//                       w e^2, p, t, p^2, t^2, p t of the (optionally affine-denormalised) prediction / target.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) normalize_kernel(float* __restrict__ x, const float* __restrict__ mean,
                                                        const float* __restrict__ stdv, const int* __restrict__ kind,
                                                        int V, long long hw) {
  const int plane = blockIdx.y;                 // b * V + v
  const int v = plane % V;
  float* p = x + (size_t)plane * hw;
  const int k = kind[v];
  const float mu = mean[v], sd = stdv[v];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    float t = p[i];
    if (k == 0) {
      t = (t - mu) / sd;                        // torchvision Normalize: sub then div (IEEE division, like the reference)
    } else {
      t *= 1000.f;                              // m -> mm
      t = (t <= 0.25f) ? 0.f : t;               // below 0.25 mm/day counts as dry
      t = log1pf(t);
    }
    p[i] = t;
  }
}

struct EvalArgs {
  const void* pred; const float* target; const float* lat_w; const float* scale; const float* shift;
  double* out; int B, C, H, W, tH, tW;
};

template <typename T>
__global__ void __launch_bounds__(256) eval_stats_kernel(const EvalArgs a) {
  const int plane = blockIdx.y;                 // b * C + c
  const int b = plane / a.C, c = plane % a.C;
  const T* p = reinterpret_cast<const T*>(a.pred) + (size_t)plane * a.H * a.W;
  const float* t = a.target + ((size_t)b * a.C + c) * a.tH * a.tW;
  const float sc = a.scale ? a.scale[c] : 1.f, sh = a.shift ? a.shift[c] : 0.f;
  double s[6] = {0, 0, 0, 0, 0, 0};
  const long long n = (long long)a.H * a.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / a.W), x = (int)(i % a.W);
    const float pv = fmaf(to_f(p[i]), sc, sh);
    const float tv = fmaf(t[(size_t)y * a.tW + x], sc, sh);
    const float e = pv - tv;
    const float w = a.lat_w ? a.lat_w[y] : 1.f;
    s[0] += (double)(w * e * e);
    s[1] += pv; s[2] += tv;
    s[3] += (double)pv * pv; s[4] += (double)tv * tv; s[5] += (double)pv * tv;
  }
  __shared__ double sm[6][8];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += sm[threadIdx.x][w];
    atomicAdd(&a.out[(size_t)plane * 6 + threadIdx.x], v);
  }
}

}  // namespace

extern "C" int o2_normalize_fields(float* x, const float* mean, const float* stdv, const int* kind, int B, int V, int64_t hw,
                                   void* stream) {
  O2_REQUIRE(x && mean && stdv && kind, "normalize_fields: null pointer");
  O2_REQUIRE(B > 0 && V > 0 && hw > 0, "normalize_fields: empty problem");
  O2_REQUIRE((long long)B * V <= 65535, "normalize_fields: B*V too large");
  long long gx = (hw + 255) / 256;
  const long long cap = ((long long)o2_num_sms() * 8 + (long long)B * V - 1) / ((long long)B * V);
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  normalize_kernel<<<dim3((unsigned)gx, (unsigned)(B * V)), 256, 0, (cudaStream_t)stream>>>(x, mean, stdv, kind, V, hw);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

extern "C" int o2_eval_stats(const void* pred, int dtype, const float* target, const float* lat_w, const float* scale,
                             const float* shift, double* out, int B, int C, int H, int W, int tgt_H, int tgt_W,
                             void* stream) {
  O2_REQUIRE(pred && target && out, "eval_stats: null pointer");
  O2_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && tgt_H >= H && tgt_W >= W, "eval_stats: bad dims");
  O2_REQUIRE((long long)B * C <= 65535, "eval_stats: B*C too large");
  EvalArgs a;
  a.pred = pred; a.target = target; a.lat_w = lat_w; a.scale = scale; a.shift = shift; a.out = out;
  a.B = B; a.C = C; a.H = H; a.W = W; a.tH = tgt_H; a.tW = tgt_W;
  cudaStream_t st = (cudaStream_t)stream;
  O2_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 6 * B * C, st));
  long long gx = ((long long)H * W + 255) / 256;
  const long long cap = ((long long)o2_num_sms() * 8 + (long long)B * C - 1) / ((long long)B * C);
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  dim3 grid((unsigned)gx, (unsigned)(B * C));
  if (dtype == O2_F32) eval_stats_kernel<float><<<grid, 256, 0, st>>>(a);
  else if (dtype == O2_BF16) eval_stats_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a);
  else O2_FAIL(O2_ERR_ARG, "eval_stats: bad dtype %d", dtype);
  O2_LAUNCH_CHECK();
  return O2_OK;
}
