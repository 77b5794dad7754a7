// Fused clip_replace_constant + {mse, mae, bayesian_tv} forward and gradient (HBM-bound, one pass).
//
// reference: examples/intermediate_downscaling.py:267-278 (clamp precip >= 0, overwrite CONSTANT channels with the
// target), metrics/functional.py:173-202 (mse), :218-232 (mae), :117-167 (bayesian_tv = mse + 0.02 * 4-direction total
// variation of the prediction), latitude weights metrics/metrics.py:58-65, channel weights functional.py:188-196.
//
// One CTA stages a (TH+2) x (TW+2) halo tile of the *clipped* prediction in shared memory (coalesced reads along W),
// then every thread produces 4 pixels: weighted error -> warp-shuffle/CTA reduction -> one fp64 atomic per CTA, and the
// analytic gradient (gather form of the TV stencil, sign(0)=0 like torch.abs) written once.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int TW = 128, TH = 8, NT = 256;

struct LossArgs {
  const void* pred; const float* target; void* dpred;
  const float* lat_w; const float* ch_w; double* accum;
  int kind, clamp_ch; uint32_t const_mask;
  int B, C, H, W, tgt_H, tgt_W;
  float gscale;  // grad_scale / (B*C*H*W)
};

template <typename T>
__global__ void __launch_bounds__(NT) loss_kernel(const LossArgs a) {
  __shared__ float sp[TH + 2][TW + 2];
  __shared__ unsigned char spass[TH][TW];
  __shared__ float swarp[NT / 32];
  const int bc = blockIdx.z;
  const int c = bc % a.C;
  const int h0 = blockIdx.y * TH, w0 = blockIdx.x * TW;
  const T* pred = reinterpret_cast<const T*>(a.pred) + (size_t)bc * a.H * a.W;
  const float* tgt = a.target + (size_t)bc * a.tgt_H * a.tgt_W;
  const bool is_const = (a.const_mask >> c) & 1u;
  const bool is_clamp = (c == a.clamp_ch);
  const bool tv = (a.kind == O2_LOSS_BAYESIAN_TV);

  for (int i = threadIdx.x; i < (TH + 2) * (TW + 2); i += NT) {
    const int r = i / (TW + 2), cc = i % (TW + 2);
    const int h = h0 + r - 1, w = w0 + cc - 1;
    float v = 0.f;
    bool pass = false;
    if (h >= 0 && h < a.H && w >= 0 && w < a.W) {
      if (is_const) {
        v = tgt[(size_t)h * a.tgt_W + w];
      } else {
        v = to_f(pred[(size_t)h * a.W + w]);
        pass = true;
        if (is_clamp && v < 0.f) { v = 0.f; pass = false; }   // clamp_(min=0): gradient only where raw >= 0
      }
    }
    sp[r][cc] = v;
    if (r >= 1 && r <= TH && cc >= 1 && cc <= TW) spass[r - 1][cc - 1] = pass ? 1 : 0;
  }
  __syncthreads();

  const float chw = a.ch_w ? a.ch_w[c] : 1.0f;
  float local = 0.f;
#pragma unroll
  for (int it = 0; it < (TH * TW) / NT; ++it) {
    const int idx = it * NT + threadIdx.x;
    const int r = idx / TW, cc = idx % TW;
    const int h = h0 + r, w = w0 + cc;
    if (h >= a.H || w >= a.W) continue;
    const float p = sp[r + 1][cc + 1];
    const float t = tgt[(size_t)h * a.tgt_W + w];
    const float lw = a.lat_w ? a.lat_w[h] : 1.0f;
    const float d = p - t;
    float err, g;
    if (a.kind == O2_LOSS_MAE) {
      err = fabsf(d);
      g = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
    } else {
      err = d * d;
      g = 2.f * d;
    }
    g *= lw;
    if (tv) {
      const bool hb = (h + 1 < a.H), ht = (h >= 1), wr = (w + 1 < a.W), wl = (w >= 1);
      const float lwm = (a.lat_w && ht) ? a.lat_w[h - 1] : 1.0f;
      auto sgn = [](float x) { return (x > 0.f) ? 1.f : (x < 0.f ? -1.f : 0.f); };
      // own cell: |p(h+1,w)-p|, |p(h,w+1)-p|, 0.7|p(h+1,w+1)-p|, 0.7|p(h+1,w-1)-p|
      float e = 0.f, gs = 0.f;
      if (hb) { const float x = sp[r + 2][cc + 1] - p; e += fabsf(x); gs -= sgn(x); }
      if (wr) { const float x = sp[r + 1][cc + 2] - p; e += fabsf(x); gs -= sgn(x); }
      if (hb && wr) { const float x = sp[r + 2][cc + 2] - p; e += 0.7f * fabsf(x); gs -= 0.7f * sgn(x); }
      if (hb && wl) { const float x = sp[r + 2][cc] - p; e += 0.7f * fabsf(x); gs -= 0.7f * sgn(x); }
      err += 0.02f * e;
      float gn = lw * gs;
      // cells that reference p as their "+" neighbour: (h, w-1) same row; (h-1, w), (h-1, w-1), (h-1, w+1) row above
      if (wl) gn += lw * sgn(p - sp[r + 1][cc]);
      if (ht) {
        float s = sgn(p - sp[r][cc + 1]);
        if (wl) s += 0.7f * sgn(p - sp[r][cc]);
        if (wr) s += 0.7f * sgn(p - sp[r][cc + 2]);
        gn += lwm * s;
      }
      g += 0.02f * gn;
    }
    local += err * lw;
    if (a.dpred) {
      const float gv = spass[r][cc] ? g * chw * a.gscale : 0.f;
      reinterpret_cast<T*>(a.dpred)[(size_t)bc * a.H * a.W + (size_t)h * a.W + w] = from_f<T>(gv);
    }
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < NT / 32; ++i) s += swarp[i];
    atomicAdd(&a.accum[c], (double)s);
  }
}


// Streaming variant for mse / mae (no shared memory, no barriers): one thread owns 8 consecutive pixels of one row, reads
// them with 16-byte loads (prediction) and two 16-byte loads (target) and writes its 8 gradients with one 16/32-byte
// store.  Used when W % 8 == 0 and the rows are 16-byte aligned; loss_kernel (tiled through shared memory) handles every
// other shape, loss_tv_band_kernel the bayesian_tv stencil.
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 t;
  t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]); t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = t;
}

template <typename T>
__global__ void __launch_bounds__(NT) loss_stream_kernel(const LossArgs a) {
  __shared__ float swarp[NT / 32];
  const int bc = blockIdx.y;
  const int c = bc % a.C;
  const int W8 = a.W >> 3;
  const long long item = (long long)blockIdx.x * NT + threadIdx.x;
  const bool active = item < (long long)a.H * W8;
  const int h = active ? (int)(item / W8) : 0;
  const int w0 = active ? (int)(item % W8) * 8 : 0;
  const T* pred = reinterpret_cast<const T*>(a.pred) + (size_t)bc * a.H * a.W;
  const float* tgt = a.target + (size_t)bc * a.tgt_H * a.tgt_W;
  const bool is_const = (a.const_mask >> c) & 1u;
  const bool is_clamp = (c == a.clamp_ch);
  float local = 0.f;
  if (active) {
    float t[8], pv[8], gout[8];
    load8<float>(tgt + (size_t)h * a.tgt_W + w0, t);
    if (is_const) {
#pragma unroll
      for (int i = 0; i < 8; ++i) pv[i] = t[i];
    } else {
      load8<T>(pred + (size_t)h * a.W + w0, pv);
    }
    const float lw = a.lat_w ? a.lat_w[h] : 1.0f;
    const float chw = a.ch_w ? a.ch_w[c] : 1.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      bool pass = !is_const;
      float p = pv[i];
      if (is_clamp && p < 0.f) { p = 0.f; pass = false; }       // clamp_(min=0): gradient only where raw >= 0
      const float d = p - t[i];
      float err, g;
      if (a.kind == O2_LOSS_MAE) { err = fabsf(d); g = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f); }
      else { err = d * d; g = 2.f * d; }
      local += err * lw;
      gout[i] = pass ? g * lw * chw * a.gscale : 0.f;
    }
    if (a.dpred) store8<T>(reinterpret_cast<T*>(a.dpred) + (size_t)bc * a.H * a.W + (size_t)h * a.W + w0, gout);
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s_ = 0.f;
    for (int i = 0; i < NT / 32; ++i) s_ += swarp[i];
    atomicAdd(&a.accum[c], (double)s_);
  }
}

// Band variant for bayesian_tv (the shipped train_loss): one thread owns an 8-column strip of RH consecutive rows and
// slides a two-row register window down it, so DRAM / L1 see every prediction row (RH + 2) / RH times instead of 3, the
// halo columns come from the neighbour lanes by warp shuffle (only lanes 0 / 31 load a scalar), and every TV difference
// that lies inside the strip is evaluated ONCE and scattered to both of its end points (35 sign evaluations per 8
// pixels instead of 64).  Pair ownership follows functional.py:141-160: pixel (h, w) owns |p(h+1,w)-p|, |p(h,w+1)-p|,
// 0.7|p(h+1,w+1)-p|, 0.7|p(h+1,w-1)-p| and weights them with lat_w[h]; a pair whose end point lies outside the image
// does not exist (the reference zero-pads the difference tensors).
constexpr int RH = 8;

// raw (not yet unpacked) CW-element row segments: the loads of the next row stay in flight while the current row computes
template <typename S, int CW> struct RawN;
template <> struct RawN<float, 8> { float4 a, b; };
template <> struct RawN<__nv_bfloat16, 8> { uint4 a; };
template <> struct RawN<float, 4> { float4 a; };
template <> struct RawN<__nv_bfloat16, 4> { uint2 a; };
__device__ __forceinline__ void ldraw(const float* p, RawN<float, 8>& r) {
  r.a = __ldg(reinterpret_cast<const float4*>(p));
  r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
}
__device__ __forceinline__ void ldraw(const __nv_bfloat16* p, RawN<__nv_bfloat16, 8>& r) {
  r.a = __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void ldraw(const float* p, RawN<float, 4>& r) { r.a = __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void ldraw(const __nv_bfloat16* p, RawN<__nv_bfloat16, 4>& r) {
  r.a = __ldg(reinterpret_cast<const uint2*>(p));
}
__device__ __forceinline__ void unpackN(const RawN<float, 8>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}
__device__ __forceinline__ void unpackN(const RawN<__nv_bfloat16, 8>& r, float (&v)[8]) {
  const float2 a = unpack_bf16x2(r.a.x), b = unpack_bf16x2(r.a.y), c = unpack_bf16x2(r.a.z), d = unpack_bf16x2(r.a.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void unpackN(const RawN<float, 4>& r, float (&v)[4]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w;
}
__device__ __forceinline__ void unpackN(const RawN<__nv_bfloat16, 4>& r, float (&v)[4]) {
  const float2 a = unpack_bf16x2(r.a.x), b = unpack_bf16x2(r.a.y);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void storeN(float* p, const float (&v)[8]) { store8<float>(p, v); }
__device__ __forceinline__ void storeN(__nv_bfloat16* p, const float (&v)[8]) { store8<__nv_bfloat16>(p, v); }
__device__ __forceinline__ void storeN(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void storeN(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 t;
  t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}

template <typename T, bool CONST, int CW>     // CW columns per thread; CONST: the channel is overwritten with the target (no gradient, the TV term is the target's)
__device__ __forceinline__ void loss_tv_band_body(const LossArgs& a, float* swarp) {
  using S = typename std::conditional<CONST, float, T>::type;       // element type of the rows the TV stencil reads
  const int bc = blockIdx.y;
  const int c = bc % a.C;
  const int W8 = a.W / CW;
  const int warps_per_row = (W8 + 31) >> 5;
  const int n_bands = (a.H + RH - 1) / RH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * (NT / 32) + warp;
  const bool warp_ok = gw < (long long)n_bands * warps_per_row;          // warp-uniform
  const int band = warp_ok ? (int)(gw / warps_per_row) : 0;
  const int grp = (warp_ok ? (int)(gw % warps_per_row) : 0) * 32 + lane;
  const bool active = warp_ok && grp < W8;
  const int w0 = active ? grp * CW : 0;                                   // idle lanes shadow group 0 (valid addresses)
  const int h0 = band * RH;
  const float* tgt = a.target + (size_t)bc * a.tgt_H * a.tgt_W + w0;
  const S* src;
  size_t sW;
  if constexpr (CONST) { src = tgt; sW = (size_t)a.tgt_W; }
  else { src = reinterpret_cast<const T*>(a.pred) + (size_t)bc * a.H * a.W + w0; sW = (size_t)a.W; }
  const bool is_clamp = (c == a.clamp_ch);
  const bool validL = w0 > 0, validR = w0 + CW < a.W;
  const bool haloL = (lane == 0) && validL, haloR = (lane == 31) && validR;
  const float chw = (a.ch_w ? a.ch_w[c] : 1.0f) * a.gscale;
  const int Hm1 = a.H - 1;
  float local = 0.f;

  // issue the loads of row hh (clamped into the image; callers mask what a clamped row feeds): CW own columns + the halo
  // columns w0 - 1 / w0 + CW that no neighbour lane holds
  auto issue = [&](int hh, RawN<S, CW>& rw, float& l, float& r) {
    hh = hh < 0 ? 0 : (hh > Hm1 ? Hm1 : hh);
    const S* rowp = src + (size_t)hh * sW;
    ldraw(rowp, rw);
    l = 0.f; r = 0.f;
    if (haloL) l = to_f(__ldg(rowp - 1));
    if (haloR) r = to_f(__ldg(rowp + CW));
  };
  auto issue_t = [&](int hh, RawN<float, CW>& rt) {
    hh = hh > Hm1 ? Hm1 : hh;
    ldraw(tgt + (size_t)hh * a.tgt_W, rt);
  };
  // landed loads -> clipped values at columns w0-1 .. w0+CW and the gradient-pass bits of the CW own columns
  auto finish = [&](const RawN<S, CW>& rw, float l, float r, float (&dst)[CW + 2], uint32_t& pass) {
    float v[CW];
    unpackN(rw, v);
    pass = CONST ? 0u : ((1u << CW) - 1u);
    if (is_clamp) {
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        if (v[i] < 0.f) { v[i] = 0.f; pass &= ~(1u << i); }
      }
      l = fmaxf(l, 0.f); r = fmaxf(r, 0.f);
    }
#pragma unroll
    for (int i = 0; i < CW; ++i) dst[i + 1] = v[i];
    const float ls = __shfl_up_sync(0xffffffffu, v[CW - 1], 1), rs = __shfl_down_sync(0xffffffffu, v[0], 1);
    dst[0] = lane == 0 ? l : ls;
    dst[CW + 1] = lane == 31 ? r : rs;
  };
  // d|x|/dx with sign(0) = 0 (torch.abs): +-1 (FSET.BF + LOP3) and +-0.7
  auto sgn1 = [](float x) {
    return __uint_as_float(__float_as_uint(x != 0.f ? 1.f : 0.f) | (__float_as_uint(x) & 0x80000000u));
  };
  auto sgn07 = [](float x) {
    return __uint_as_float(__float_as_uint(x != 0.f ? 0.7f : 0.f) | (__float_as_uint(x) & 0x80000000u));
  };

  if (warp_ok) {
    const bool full = (h0 + RH < a.H);                   // every row of the band has a row below it (warp-uniform)
    // lat_w[h0 - 1 + lane] (lanes 0 .. RH), broadcast by shuffle where a row needs it: no load on the per-row path
    float lw_lane = 1.0f;
    if (a.lat_w && lane <= RH) {
      int hh = h0 - 1 + lane;
      hh = hh < 0 ? 0 : (hh > Hm1 ? Hm1 : hh);
      lw_lane = __ldg(a.lat_w + hh);
    }
    // pull the whole band towards L2 first: the register pipeline below runs one row ahead, which covers an L2 hit
    // (~0.4 us) but not a DRAM miss under load (> 1 us)
#pragma unroll
    for (int k = 1; k <= RH + 2; ++k) {
      int hh = h0 - 1 + k;
      hh = hh > Hm1 ? Hm1 : hh;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)hh * sW));
      if (!CONST && k <= RH) asm volatile("prefetch.global.L2 [%0];" ::"l"(tgt + (size_t)hh * a.tgt_W));
    }
    float cur[CW + 2], nxt[CW + 2], carry[CW];
    uint32_t pc, pn;
    RawN<S, CW> rp, rp2;
    RawN<float, CW> rt;
    float rl, rr, rl2, rr2;
    issue(h0 - 1, rp, rl, rr);
    issue(h0, rp2, rl2, rr2);
    finish(rp, rl, rr, cur, pn);
    issue(h0 + 1, rp, rl, rr);                           // stays in flight until the first row step
    if (!CONST) issue_t(h0, rt);
    finish(rp2, rl2, rr2, nxt, pc);
#pragma unroll
    for (int i = 0; i < CW; ++i) carry[i] = 0.f;
    if (h0 > 0) {
      // what the pairs owned by row h0 - 1 (another band's row) hand down to row h0
      const float lwm = __shfl_sync(0xffffffffu, lw_lane, 0);
#pragma unroll
      for (int i = 0; i < CW; ++i) carry[i] = sgn1(nxt[i + 1] - cur[i + 1]);
#pragma unroll
      for (int j = 0; j <= CW - 1; ++j) {
        float x = nxt[j + 1] - cur[j];
        if (j == 0 && !validL) x = 0.f;
        carry[j] += sgn07(x);
      }
#pragma unroll
      for (int j = 2; j <= CW + 1; ++j) {
        float x = nxt[j - 1] - cur[j];
        if (j == CW + 1 && !validR) x = 0.f;
        carry[j - 2] += sgn07(x);
      }
#pragma unroll
      for (int i = 0; i < CW; ++i) carry[i] *= lwm;
    }
#pragma unroll
    for (int i = 0; i < CW + 2; ++i) cur[i] = nxt[i];

    // one row: cur = row h, rp = loads of row h + 1 (issued one step ago), rt = loads of the target of row h
    auto row_step = [&](int h, bool below) {
      finish(rp, rl, rr, nxt, pn);
      float t[CW];
      if (!CONST) unpackN(rt, t);
      issue(h + 2, rp, rl, rr);                          // in flight while this row computes
      if (!CONST) issue_t(h + 1, rt);
      const float lw = __shfl_sync(0xffffffffu, lw_lane, h - h0 + 1);
      float gs[CW], gn[CW];
      float eA = 0.f, eB = 0.f;
      {  // |p(h, w+1) - p(h, w)|: owner array index j (column w0 - 1 + j), other j + 1
        float x = cur[1] - cur[0];
        if (!validL) x = 0.f;
        gs[0] = sgn1(x);
#pragma unroll
        for (int j = 1; j <= CW; ++j) {
          x = cur[j + 1] - cur[j];
          if (j == CW && !validR) x = 0.f;
          const float s = sgn1(x);
          eA += fabsf(x); gs[j - 1] -= s;
          if (j <= CW - 1) gs[j] = s;
        }
      }
#pragma unroll
      for (int i = 0; i < CW; ++i) gn[i] = 0.f;
      if (below) {
#pragma unroll
        for (int i = 0; i < CW; ++i) {                    // |p(h+1, w) - p(h, w)|
          const float x = nxt[i + 1] - cur[i + 1];
          const float s = sgn1(x);
          eA += fabsf(x); gs[i] -= s; gn[i] = s;
        }
#pragma unroll
        for (int j = 0; j <= CW; ++j) {                   // 0.7 |p(h+1, w+1) - p(h, w)|: owner j, other j + 1 in the row below
          float x = nxt[j + 1] - cur[j];
          if (j == 0 && !validL) x = 0.f;
          if (j == CW && !validR) x = 0.f;
          const float s = sgn07(x);
          if (j >= 1) { eB += fabsf(x); gs[j - 1] -= s; }
          if (j <= CW - 1) gn[j] += s;
        }
#pragma unroll
        for (int j = 1; j <= CW + 1; ++j) {                   // 0.7 |p(h+1, w-1) - p(h, w)|: owner j, other j - 1 in the row below
          float x = nxt[j - 1] - cur[j];
          if (j == 1 && !validL) x = 0.f;
          if (j == CW + 1 && !validR) x = 0.f;
          const float s = sgn07(x);
          if (j <= CW) { eB += fabsf(x); gs[j - 1] -= s; }
          if (j >= 2) gn[j - 2] += s;
        }
      }
      float gout[CW], e2 = 0.f;
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        const float d = CONST ? 0.f : cur[i + 1] - t[i];
        e2 = fmaf(d, d, e2);
        const float g = 2.f * d * lw + 0.02f * (lw * gs[i] + carry[i]);
        gout[i] = ((pc >> i) & 1u) ? g * chw : 0.f;
      }
      if (active) {
        local += lw * (e2 + 0.02f * (eA + 0.7f * eB));
        if (a.dpred) storeN(reinterpret_cast<T*>(a.dpred) + (size_t)bc * a.H * a.W + (size_t)h * a.W + w0, gout);
      }
#pragma unroll
      for (int i = 0; i < CW; ++i) carry[i] = lw * gn[i];
#pragma unroll
      for (int i = 0; i < CW + 2; ++i) cur[i] = nxt[i];
      pc = pn;
    };
    if (full) {
#pragma unroll
      for (int k = 0; k < RH; ++k) row_step(h0 + k, true);
    } else {
#pragma unroll 1
      for (int h = h0; h < a.H; ++h) row_step(h, h + 1 < a.H);
    }
  }
  local = warp_sum(local);
  if (lane == 0) swarp[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s_ = 0.f;
    for (int i = 0; i < NT / 32; ++i) s_ += swarp[i];
    atomicAdd(&a.accum[c], (double)s_);
  }
}

template <typename T, int CW>
__global__ void __launch_bounds__(NT, CW == 4 ? 3 : 2) loss_tv_band_kernel(const LossArgs a) {
  __shared__ float swarp[NT / 32];
  if ((a.const_mask >> (blockIdx.y % a.C)) & 1u) loss_tv_band_body<T, true, CW>(a, swarp);     // CTA-uniform
  else loss_tv_band_body<T, false, CW>(a, swarp);
}

__global__ void loss_finalize(const double* accum, const float* ch_w, float* loss_vec, int C, double inv_per_ch) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double tot = 0.0;
    for (int c = 0; c < C; ++c) {
      const double v = accum[c] * (ch_w ? (double)ch_w[c] : 1.0) * inv_per_ch;
      loss_vec[c] = (float)v;
      tot += v;
    }
    loss_vec[C] = (float)(tot / C);
  }
}

template <typename T>
__global__ void clip_replace_kernel(T* pred, const float* target, int clamp_ch, uint32_t const_mask, int C, int H, int W,
                                    int tgt_H, int tgt_W, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const size_t bc = i / ((size_t)W * H);
    const int c = (int)(bc % C);
    if ((const_mask >> c) & 1u) pred[i] = from_f<T>(target[bc * tgt_H * tgt_W + (size_t)h * tgt_W + w]);
    else if (c == clamp_ch) { const float v = to_f(pred[i]); if (v < 0.f) pred[i] = from_f<T>(0.f); }
  }
}

template <typename T>
__global__ void scale_channels_kernel(T* g, const float* __restrict__ scale, int C, size_t hw, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / hw) % C);
    g[i] = from_f<T>(to_f(g[i]) * scale[c]);
  }
}

}  // namespace

extern "C" int o2_scale_channels(void* g, int dtype, const float* scale, int B, int C, int64_t hw, void* stream) {
  O2_REQUIRE(g && scale && B > 0 && C > 0 && hw > 0, "scale_channels: bad args");
  const size_t total = (size_t)B * C * hw;
  const size_t want = (total + 255) / 256, cap = (size_t)o2_num_sms() * 8;
  const int grid = (int)(want < cap ? want : cap);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == O2_F32) scale_channels_kernel<float><<<grid, 256, 0, st>>>((float*)g, scale, C, (size_t)hw, total);
  else if (dtype == O2_BF16)
    scale_channels_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)g, scale, C, (size_t)hw, total);
  else O2_FAIL(O2_ERR_ARG, "scale_channels: bad dtype %d", dtype);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

extern "C" int o2_loss_fwd_bwd(const void* pred, int dtype, const float* target, void* dpred, float* loss_vec,
                               double* accum_ws, const float* lat_w, const float* ch_w, int kind, int clamp_ch,
                               uint32_t const_mask, int B, int C, int H, int W, int tgt_H, int tgt_W, float grad_scale,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  O2_REQUIRE(pred && target && loss_vec && accum_ws, "loss: null pointer");
  O2_REQUIRE(B > 0 && C > 0 && C <= 32 && H > 0 && W > 0, "loss: bad dims B=%d C=%d H=%d W=%d", B, C, H, W);
  O2_REQUIRE(tgt_H >= H && tgt_W >= W, "loss: target %dx%d smaller than prediction %dx%d", tgt_H, tgt_W, H, W);
  O2_REQUIRE(kind >= O2_LOSS_MSE && kind <= O2_LOSS_BAYESIAN_TV, "loss: unknown kind %d", kind);
  O2_REQUIRE(dtype == O2_F32 || dtype == O2_BF16, "loss: bad dtype %d", dtype);
  O2_REQUIRE((long long)B * C <= 65535, "loss: B*C too large");
  LossArgs a;
  a.pred = pred; a.target = target; a.dpred = dpred; a.lat_w = lat_w; a.ch_w = ch_w; a.accum = accum_ws;
  a.kind = kind; a.clamp_ch = clamp_ch; a.const_mask = const_mask;
  a.B = B; a.C = C; a.H = H; a.W = W; a.tgt_H = tgt_H; a.tgt_W = tgt_W;
  a.gscale = (float)((double)grad_scale / ((double)B * C * H * W));
  O2_CUDA(cudaMemsetAsync(accum_ws, 0, sizeof(double) * C, st));
  const size_t esz = dtype == O2_F32 ? 4 : 2;
  const bool stream_ok = (W % 8 == 0) && (tgt_W % 4 == 0) && ((uintptr_t)pred % 16 == 0) && ((uintptr_t)target % 16 == 0) &&
                         (dpred == nullptr || (uintptr_t)dpred % 16 == 0) && (((size_t)H * W * esz) % 16 == 0) &&
                         (((size_t)tgt_H * tgt_W * 4) % 16 == 0);
  if (stream_ok) {
    if (kind == O2_LOSS_BAYESIAN_TV) {
      static const int cw = getenv("O2_LOSS_CW") ? atoi(getenv("O2_LOSS_CW")) : 8;     // columns per thread (A/B switch: 4 | 8)
      const long long warps = (long long)((H + RH - 1) / RH) * ((W / cw + 31) / 32);
      dim3 grid((unsigned)((warps + NT / 32 - 1) / (NT / 32)), B * C);
      if (cw == 8) {
        if (dtype == O2_F32) loss_tv_band_kernel<float, 8><<<grid, NT, 0, st>>>(a);
        else loss_tv_band_kernel<__nv_bfloat16, 8><<<grid, NT, 0, st>>>(a);
      } else {
        if (dtype == O2_F32) loss_tv_band_kernel<float, 4><<<grid, NT, 0, st>>>(a);
        else loss_tv_band_kernel<__nv_bfloat16, 4><<<grid, NT, 0, st>>>(a);
      }
    } else {
      const long long items = (long long)H * (W / 8);
      dim3 grid((unsigned)((items + NT - 1) / NT), B * C);
      if (dtype == O2_F32) loss_stream_kernel<float><<<grid, NT, 0, st>>>(a);
      else loss_stream_kernel<__nv_bfloat16><<<grid, NT, 0, st>>>(a);
    }
  } else {
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B * C);
    if (dtype == O2_F32) loss_kernel<float><<<grid, NT, 0, st>>>(a);
    else loss_kernel<__nv_bfloat16><<<grid, NT, 0, st>>>(a);
  }
  O2_LAUNCH_CHECK();
  loss_finalize<<<1, 32, 0, st>>>(accum_ws, ch_w, loss_vec, C, 1.0 / ((double)B * H * W));
  O2_LAUNCH_CHECK();
  return O2_OK;
}

extern "C" int o2_clip_replace(void* pred, int dtype, const float* target, int clamp_ch, uint32_t const_mask, int B, int C,
                               int H, int W, int tgt_H, int tgt_W, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  O2_REQUIRE(pred && target, "clip_replace: null pointer");
  O2_REQUIRE(tgt_H >= H && tgt_W >= W, "clip_replace: target smaller than prediction");
  const size_t total = (size_t)B * C * H * W;
  const int grid = (int)((total + 255) / 256 < (size_t)o2_num_sms() * 8 ? (total + 255) / 256 : (size_t)o2_num_sms() * 8);
  if (dtype == O2_F32)
    clip_replace_kernel<float><<<grid, 256, 0, st>>>((float*)pred, target, clamp_ch, const_mask, C, H, W, tgt_H, tgt_W, total);
  else if (dtype == O2_BF16)
    clip_replace_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)pred, target, clamp_ch, const_mask, C, H, W, tgt_H, tgt_W, total);
  else O2_FAIL(O2_ERR_ARG, "clip_replace: bad dtype %d", dtype);
  O2_LAUNCH_CHECK();
  return O2_OK;
}
