// Small HBM-bound helpers: fp32->bf16 cast (compute copies of the fp32 master weights), column sums (bias
// gradients), fused AdamW (torch.optim.AdamW semantics; reference optimizer: intermediate_downscaling.py:642-644).
#include "common.cuh"

namespace {

__global__ void cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[n4 * 4 + threadIdx.x] = __float2bfloat16_rn(src[n4 * 4 + threadIdx.x]);
}

template <typename T, int VN>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, float* __restrict__ out, long long M,
                                                     long long N, long long ld) {
  __shared__ float sm[8][32 * VN + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long col = ((long long)blockIdx.x * 32 + tx) * VN;
  float acc[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) acc[j] = 0.f;
  if (col < N) {
    for (long long m = (long long)blockIdx.y * 8 + ty; m < M; m += (long long)gridDim.y * 8) {
      if constexpr (VN == 8) {
        const uint4 t = *reinterpret_cast<const uint4*>(X + m * ld + col);
        const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
        acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
      } else {
        const float4 t = *reinterpret_cast<const float4*>(X + m * ld + col);
        acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VN; ++j) sm[ty][tx * VN + j] = acc[j];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * VN; i += 256) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += sm[r][i];
    const long long c = (long long)blockIdx.x * 32 * VN + i;
    if (c < N) atomicAdd(&out[c], s);
  }
}

__device__ __forceinline__ float adamw_one(float& pi, float gi, float& mi, float& vi, float lr, float b1, float b2,
                                          float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  gi *= gscale;
  pi *= (1.f - lr * wd);                       // decoupled weight decay (torch.optim.AdamW)
  mi = b1 * mi + (1.f - b1) * gi;
  vi = b2 * vi + (1.f - b2) * gi * gi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi -= (lr / bc1) * (mi / denom);
  return pi;
}

// 16-byte vectors when every buffer is 16-byte aligned (the flat layout aligns every slice to 8 elements), scalar tail
__device__ __forceinline__ void adamw_update(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                             float* __restrict__ v, __nv_bfloat16* __restrict__ pb, long long n, float lr,
                                             float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                             float gscale) {
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)pb & 7) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    adamw_one(p4.x, g4.x, m4.x, v4.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    adamw_one(p4.y, g4.y, m4.y, v4.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    adamw_one(p4.z, g4.z, m4.z, v4.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    adamw_one(p4.w, g4.w, m4.w, v4.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    reinterpret_cast<float4*>(p)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    if (pb) {
      uint2 t;
      t.x = pack_bf16x2(p4.x, p4.y); t.y = pack_bf16x2(p4.z, p4.w);
      reinterpret_cast<uint2*>(pb)[i] = t;
    }
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    adamw_one(pi, g[i], mi, vi, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (pb) pb[i] = __float2bfloat16_rn(pi);
  }
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, __nv_bfloat16* __restrict__ pb, long long n, float lr, float b1,
                             float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  adamw_update(p, g, m, v, pb, n, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
}

// same update with the step-dependent scalars in device memory: a captured CUDA graph of the step stays valid while the
// host rewrites {lr, beta1, beta2, eps, weight_decay, 1 - beta1^t, sqrt(1 - beta2^t), grad_scale} between replays
__global__ void adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, __nv_bfloat16* __restrict__ pb, long long n,
                                 const float* __restrict__ sc) {
  adamw_update(p, g, m, v, pb, n, sc[0], sc[1], sc[2], sc[3], sc[4], sc[5], sc[6], sc[7]);
}

inline int grid_for(long long n, int per_block) {
  long long want = (n + per_block - 1) / per_block;
  long long cap = (long long)o2_num_sms() * 8;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace

__global__ void uncast_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint2 t = reinterpret_cast<const uint2*>(src)[i];
    const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
    reinterpret_cast<float4*>(dst)[i] = make_float4(a.x, a.y, b.x, b.y);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[n4 * 4 + threadIdx.x] = __bfloat162float(src[n4 * 4 + threadIdx.x]);
}

/* bf16 -> fp32 (operands of the fp32 attention arm when the bf16 path has no tensor-core kernel for a head dim) */
extern "C" int o2_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream) {
  O2_REQUIRE(src && dst && n >= 0, "uncast: bad args");
  if (n == 0) return O2_OK;
  O2_REQUIRE(((uintptr_t)src % 8) == 0 && ((uintptr_t)dst % 16) == 0, "uncast: misaligned pointers");
  uncast_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, dst, n);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

extern "C" int o2_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  O2_REQUIRE(src && dst && n >= 0, "cast: bad args");
  if (n == 0) return O2_OK;
  O2_REQUIRE(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 8) == 0, "cast: misaligned pointers");
  cast_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

extern "C" int o2_colsum(const void* X, int dtype, float* out, int64_t M, int64_t N, int64_t ld, void* stream) {
  O2_REQUIRE(X && out && M > 0 && N > 0, "colsum: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const long long cap = (long long)o2_num_sms() * 8;
  if (dtype == O2_BF16) {
    O2_REQUIRE(N % 8 == 0 && ld % 8 == 0, "colsum: N and ld must be multiples of 8 for bf16");
    const unsigned gx = (unsigned)((N + 255) / 256);
    long long gy = cap / gx; if (gy < 1) gy = 1; if (gy > (M + 7) / 8) gy = (M + 7) / 8;
    colsum_kernel<__nv_bfloat16, 8><<<dim3(gx, (unsigned)gy), 256, 0, st>>>((const __nv_bfloat16*)X, out, M, N, ld);
  } else if (dtype == O2_F32) {
    O2_REQUIRE(N % 4 == 0 && ld % 4 == 0, "colsum: N and ld must be multiples of 4 for fp32");
    const unsigned gx = (unsigned)((N + 127) / 128);
    long long gy = cap / gx; if (gy < 1) gy = 1; if (gy > (M + 7) / 8) gy = (M + 7) / 8;
    colsum_kernel<float, 4><<<dim3(gx, (unsigned)gy), 256, 0, st>>>((const float*)X, out, M, N, ld);
  } else O2_FAIL(O2_ERR_ARG, "colsum: bad dtype %d", dtype);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

extern "C" int o2_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  O2_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "adamw: bad args");
  if (n == 0) return O2_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<<<grid_for(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (__nv_bfloat16*)p_bf16, n, lr, beta1,
                                                                       beta2, eps, weight_decay, bc1, bc2_sqrt,
                                                                       grad_scale);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

extern "C" int o2_adamw_dev(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, const float* scalars,
                            void* stream) {
  O2_REQUIRE(p && g && m && v && scalars && n >= 0, "adamw_dev: bad args");
  if (n == 0) return O2_OK;
  adamw_dev_kernel<<<grid_for(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (__nv_bfloat16*)p_bf16, n, scalars);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

namespace {
__device__ __forceinline__ bool nonfinite_bits(float x) {      // exponent all ones: inf or NaN (integer test: immune to fast-math folding)
  return (__float_as_uint(x) & 0x7f800000u) == 0x7f800000u;
}
__global__ void nonfinite_kernel(const float* __restrict__ g, long long n, int* __restrict__ flag) {
  bool bad = false;
  const long long n4 = (((uintptr_t)g & 15) == 0) ? n / 4 : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    bad |= nonfinite_bits(v.x) | nonfinite_bits(v.y) | nonfinite_bits(v.z) | nonfinite_bits(v.w);
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    bad |= nonfinite_bits(g[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
}  // namespace

/* flag |= 1 if any of g[0..n) is inf or NaN (the found_inf test of torch's GradScaler.step, intermediate_downscaling.py:733-742);
 * the caller zero-fills flag.  One streaming pass, HBM-bound. */
extern "C" int o2_nonfinite(const float* g, int64_t n, int* flag, void* stream) {
  O2_REQUIRE(g && flag && n >= 0, "nonfinite: bad args");
  if (n == 0) return O2_OK;
  nonfinite_kernel<<<grid_for(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(g, n, flag);
  O2_LAUNCH_CHECK();
  return O2_OK;
}
