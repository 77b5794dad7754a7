// Dropout / stochastic depth on the token stream (HBM-bound, one pass, 16-byte vectors).
//
// reference: nn.Dropout at res_slimvit.py:284 (pos_drop), components/attention.py:81 (proj_drop), components/mlp.py:65,68
// (drop1 / drop2) and timm DropPath at components/vit_blocks.py:78-79 (per-sample keep / keep_prob).  One kernel serves
// all of them, forward and backward (the backward of x -> x * m is the same multiplication applied to the gradient):
//
//     out[r, c] = res[r, c] + y[r, c] * keep(e) / (1 - p) * sample_scale[r / rows_per_sample]        e = r * cols + c
//
// res == nullptr: no residual; sample_scale == nullptr: 1; p == 0: keep = 1.  out may alias y or res.
//
// The keep decision is a counter-based hash of the element index, so a mask is never stored: backward regenerates it
// from (seed, site).  Two elements share one 32-bit hash (16 bits each):
//     key  = lowbias32(seed_lo ^ lowbias32(site ^ seed_hi))
//     h    = lowbias32(lo32(e >> 1) ^ key ^ (hi32(e >> 1) * 0x9E3779B1))
//     keep = ((e & 1) ? h >> 16 : h & 0xFFFF) >= floor(p * 65536)
// (oracle/dropout_mask.py restates this in torch for the parity tests).
#include "common.cuh"

namespace {

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x21f0aaadu;
  x ^= x >> 15; x *= 0x735a2d97u;
  x ^= x >> 15;
  return x;
}

struct DropArgs {
  const void* y; const void* res; void* out; const float* sample_scale;
  long long n_vec, elems_per_sample;
  uint32_t key, thr16;
  const uint64_t* step_word;
  float inv_keep;
};

template <typename T> struct V;
template <> struct V<float> { static constexpr int N = 4; };
template <> struct V<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const DropArgs a) {
  constexpr int VN = V<T>::N;
  const uint32_t key = a.key ^ ptx::step_word_mix(a.step_word);
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < a.n_vec; v += (long long)gridDim.x * blockDim.x) {
    const long long e0 = v * VN;
    float s = a.inv_keep;
    if (a.sample_scale) s *= __ldg(a.sample_scale + e0 / a.elems_per_sample);
    float yv[VN], rv[VN];
    if constexpr (VN == 8) {
      const uint4 t = reinterpret_cast<const uint4*>(a.y)[v];
      const float2 p0 = unpack_bf16x2(t.x), p1 = unpack_bf16x2(t.y), p2 = unpack_bf16x2(t.z), p3 = unpack_bf16x2(t.w);
      yv[0] = p0.x; yv[1] = p0.y; yv[2] = p1.x; yv[3] = p1.y; yv[4] = p2.x; yv[5] = p2.y; yv[6] = p3.x; yv[7] = p3.y;
      if (a.res) {
        const uint4 u = reinterpret_cast<const uint4*>(a.res)[v];
        const float2 q0 = unpack_bf16x2(u.x), q1 = unpack_bf16x2(u.y), q2 = unpack_bf16x2(u.z), q3 = unpack_bf16x2(u.w);
        rv[0] = q0.x; rv[1] = q0.y; rv[2] = q1.x; rv[3] = q1.y; rv[4] = q2.x; rv[5] = q2.y; rv[6] = q3.x; rv[7] = q3.y;
      }
    } else {
      const float4 t = reinterpret_cast<const float4*>(a.y)[v];
      yv[0] = t.x; yv[1] = t.y; yv[2] = t.z; yv[3] = t.w;
      if (a.res) {
        const float4 u = reinterpret_cast<const float4*>(a.res)[v];
        rv[0] = u.x; rv[1] = u.y; rv[2] = u.z; rv[3] = u.w;
      }
    }
    if (!a.res) {
#pragma unroll
      for (int j = 0; j < VN; ++j) rv[j] = 0.f;
    }
    float o[VN];
#pragma unroll
    for (int j = 0; j < VN; j += 2) {
      const unsigned long long pair = (unsigned long long)(e0 + j) >> 1;
      const uint32_t h = lowbias32((uint32_t)pair ^ key ^ ((uint32_t)(pair >> 32) * 0x9E3779B1u));
      o[j] = rv[j] + (((h & 0xFFFFu) >= a.thr16) ? yv[j] * s : 0.f);
      o[j + 1] = rv[j + 1] + (((h >> 16) >= a.thr16) ? yv[j + 1] * s : 0.f);
    }
    if constexpr (VN == 8) {
      uint4 w;
      w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]); w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
      reinterpret_cast<uint4*>(a.out)[v] = w;
    } else {
      reinterpret_cast<float4*>(a.out)[v] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace

extern "C" int o2_dropout(const void* y, const void* res, void* out, int dtype, int64_t rows, int64_t cols,
                          int64_t rows_per_sample, float p, const float* sample_scale, uint64_t seed, uint32_t site,
                          void* stream) {
  O2_REQUIRE(y && out, "dropout: null pointer");
  O2_REQUIRE(rows > 0 && cols > 0, "dropout: empty problem");
  O2_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%f outside [0, 1)", (double)p);
  O2_REQUIRE(dtype == O2_F32 || dtype == O2_BF16, "dropout: bad dtype %d", dtype);
  const int vn = dtype == O2_F32 ? 4 : 8;
  O2_REQUIRE(cols % vn == 0, "dropout: cols=%lld must be a multiple of %d", (long long)cols, vn);
  O2_REQUIRE(((uintptr_t)y % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)res % 16) == 0,
             "dropout: pointers must be 16-byte aligned");
  O2_REQUIRE(!sample_scale || rows_per_sample > 0, "dropout: rows_per_sample must be > 0 with sample_scale");
  DropArgs a;
  a.y = y; a.res = res; a.out = out; a.sample_scale = sample_scale;
  a.n_vec = rows * cols / vn;
  a.elems_per_sample = sample_scale ? rows_per_sample * cols : 1;
  a.key = lowbias32((uint32_t)seed ^ lowbias32(site ^ (uint32_t)(seed >> 32)));
  a.step_word = o2_step_word();
  a.thr16 = (uint32_t)floor((double)p * 65536.0);
  a.inv_keep = 1.f / (1.f - p);
  long long blocks = (a.n_vec + 255) / 256;
  const long long cap = (long long)o2_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == O2_F32) dropout_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(a);
  else dropout_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}
