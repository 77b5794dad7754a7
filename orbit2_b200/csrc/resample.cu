// Bicubic resampling of the learned position embedding to the token grid of the current input (TILES mode and inference
// on grids other than the one the model was built for), forward and backward, channels-last and in place of the
// reference's permute -> F.interpolate(mode="bicubic", align_corners=False) -> permute round trip.
//
// reference: components/pos_embed.py:103-138 (interpolate_pos_embed_on_the_fly; called from res_slimvit.py:271-278 each
// forward) -- torch's upsample_bicubic2d: src = (dst + 0.5) * (in / out) - 0.5 (NOT clamped below 0), taps floor(src) - 1 ..
// + 2 clamped to [0, in - 1], cubic-convolution weights with A = -0.75, all in fp32.
//
// Layout: src [ih * iw, D], dst [oh * ow, D], fp32, D % 4 == 0 (one float4 of channels per thread).  HBM-bound and small
// (at most L x D = 16200 x 1024 floats); the backward is a GATHER over the output pixels that touch an input pixel (the
// per-axis tap weights of those candidates are rebuilt in shared memory by the block), so it is deterministic -- no atomics.
#include "common.cuh"

namespace {

constexpr float kA = -0.75f;

__device__ __forceinline__ float cc1(float x) { return ((kA + 2.f) * x - (kA + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cc2(float x) { return ((kA * x - 5.f * kA) * x + 8.f * kA) * x - 4.f * kA; }

// the four taps of output index o on an axis of n_in inputs: clamped indices and weights
__device__ __forceinline__ void taps(int o, float scale, int n_in, int (&idx)[4], float (&w)[4]) {
  const float s = scale * ((float)o + 0.5f) - 0.5f;
  const float fl = floorf(s);
  const float t = s - fl;
  const int i0 = (int)fl;
  w[0] = cc2(t + 1.f);
  w[1] = cc1(t);
  w[2] = cc1(1.f - t);
  w[3] = cc2(2.f - t);
#pragma unroll
  for (int k = 0; k < 4; ++k) idx[k] = min(max(i0 - 1 + k, 0), n_in - 1);
}

struct RsArgs {
  const float* src;
  float* dst;
  int ih, iw, oh, ow, d4;   // d4 = D / 4
  float sy, sx;             // in / out
};

__global__ void __launch_bounds__(256) bicubic_fwd_kernel(const RsArgs a) {
  const long long n = (long long)a.oh * a.ow * a.d4;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % a.d4);
    const int pix = (int)(e / a.d4);
    const int oy = pix / a.ow, ox = pix % a.ow;
    int iy[4], ix[4];
    float wy[4], wx[4];
    taps(oy, a.sy, a.ih, iy, wy);
    taps(ox, a.sx, a.iw, ix, wx);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.src) + ((long long)iy[r] * a.iw + ix[k]) * a.d4 + c);
        row.x += wx[k] * v.x; row.y += wx[k] * v.y; row.z += wx[k] * v.z; row.w += wx[k] * v.w;
      }
      acc.x += wy[r] * row.x; acc.y += wy[r] * row.y; acc.z += wy[r] * row.z; acc.w += wy[r] * row.w;
    }
    reinterpret_cast<float4*>(a.dst)[e] = acc;
  }
}

constexpr int kMaxCand = 256;   // candidate outputs per axis that may touch one input index

// first / last output index whose taps may reach input index i (over-inclusive; every candidate is re-checked)
__device__ __forceinline__ void cand_range(int i, float scale, int n_in, int n_out, int& lo, int& hi) {
  lo = (i == 0) ? 0 : max(0, (int)floorf(((float)i - 2.5f) / scale - 0.5f) - 1);
  hi = (i == n_in - 1) ? n_out - 1 : min(n_out - 1, (int)ceilf(((float)i + 2.5f) / scale - 0.5f) + 1);
}

// one block per INPUT pixel: d_src[iy, ix, :] = sum over output pixels (oy, ox) of Wy[oy, iy] Wx[ox, ix] d_dst[oy, ox, :]
__global__ void __launch_bounds__(256) bicubic_bwd_kernel(const RsArgs a) {   // here src = d_dst [oh*ow, D], dst = d_src
  __shared__ float wys[kMaxCand], wxs[kMaxCand];
  const int iy = blockIdx.x / a.iw, ix = blockIdx.x % a.iw;
  int ylo, yhi, xlo, xhi;
  cand_range(iy, a.sy, a.ih, a.oh, ylo, yhi);
  cand_range(ix, a.sx, a.iw, a.ow, xlo, xhi);
  for (int t = threadIdx.x; t <= yhi - ylo; t += blockDim.x) {
    int id[4];
    float w[4];
    taps(ylo + t, a.sy, a.ih, id, w);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += (id[k] == iy) ? w[k] : 0.f;
    wys[t] = s;
  }
  for (int t = threadIdx.x; t <= xhi - xlo; t += blockDim.x) {
    int id[4];
    float w[4];
    taps(xlo + t, a.sx, a.iw, id, w);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += (id[k] == ix) ? w[k] : 0.f;
    wxs[t] = s;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.d4; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int yy = 0; yy <= yhi - ylo; ++yy) {
      const float wy = wys[yy];
      if (wy == 0.f) continue;
      float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int xx = 0; xx <= xhi - xlo; ++xx) {
        const float wx = wxs[xx];
        if (wx == 0.f) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.src) + ((long long)(ylo + yy) * a.ow + (xlo + xx)) * a.d4 + c);
        row.x += wx * v.x; row.y += wx * v.y; row.z += wx * v.z; row.w += wx * v.w;
      }
      acc.x += wy * row.x; acc.y += wy * row.y; acc.z += wy * row.z; acc.w += wy * row.w;
    }
    reinterpret_cast<float4*>(a.dst)[(long long)blockIdx.x * a.d4 + c] = acc;
  }
}

int check(const void* src, const void* dst, int ih, int iw, int oh, int ow, int D) {
  O2_REQUIRE(src && dst, "bicubic: null pointer");
  O2_REQUIRE(ih > 0 && iw > 0 && oh > 0 && ow > 0 && D > 0, "bicubic: empty problem %dx%d -> %dx%d, D=%d", ih, iw, oh, ow, D);
  O2_REQUIRE(D % 4 == 0, "bicubic: D=%d must be a multiple of 4", D);
  O2_REQUIRE(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0, "bicubic: pointers must be 16-byte aligned");
  return O2_OK;
}

}  // namespace

extern "C" int o2_bicubic_fwd(const float* src, float* dst, int ih, int iw, int oh, int ow, int D, void* stream) {
  if (int rc = check(src, dst, ih, iw, oh, ow, D)) return rc;
  RsArgs a{src, dst, ih, iw, oh, ow, D / 4, (float)ih / (float)oh, (float)iw / (float)ow};
  const long long n = (long long)oh * ow * a.d4;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)o2_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  bicubic_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

extern "C" int o2_bicubic_bwd(const float* d_dst, float* d_src, int ih, int iw, int oh, int ow, int D, void* stream) {
  if (int rc = check(d_dst, d_src, ih, iw, oh, ow, D)) return rc;
  const float sy = (float)ih / (float)oh, sx = (float)iw / (float)ow;
  // candidates per axis: about 5 / scale + 4 outputs reach one input index (more only at the clamped borders)
  O2_REQUIRE(5.f / sy + 8.f <= kMaxCand && 5.f / sx + 8.f <= kMaxCand,
             "bicubic_bwd: magnification %dx%d -> %dx%d exceeds the %d candidate outputs per input index", ih, iw, oh, ow,
             kMaxCand);
  RsArgs a{d_dst, d_src, ih, iw, oh, ow, D / 4, sy, sx};
  bicubic_bwd_kernel<<<(unsigned)(ih * iw), 256, 0, (cudaStream_t)stream>>>(a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}
