// Upsampling head tail + residual conv branch (HBM-bound, shared-memory tiled direct convolutions).
//
// reference: res_slimvit.py:107-112 (path2 = conv3x3(C+4 -> cr*mag^2) -> GELU -> PixelShuffle(mag) -> conv3x3(cr -> C)),
// :167-179 (unpatchify = flat re-interpretation of the head output as [Ho/p][Wo/p][p][p][C]), :331 conv_out 3x3,
// :333-336 crop-add.  o2_path2_conv1_* works on the low-res grid and keeps the PRE-activation h1; o2_headtail_* fuses
// everything on the high-res grid (gather-unpatchify + conv_out + GELU/pixel-shuffle gather + conv2 + add) so that the
// prediction is written exactly once and neither the unpatchified image nor the path2 output ever exist in HBM.
#include "headtail.cuh"

using namespace o2ht;

namespace {

constexpr int TX = 32, TY = 8, NT = 256;
constexpr int HXW = TX + 2, HYW = TY + 2;

// ------------------------------------------------------------------ path2 conv1 (low-res)
template <typename T>
__global__ void __launch_bounds__(NT) conv1_fwd_kernel(const float* __restrict__ x, IdxList idx, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, T* __restrict__ h1, int B, int V,
                                                       int Hx, int Wx, int cin, int c1) {
  extern __shared__ float smem[];
  float* sx = smem;                        // [cin][HYW][HXW]
  float* sw = sx + cin * HYW * HXW;        // [c1][cin*9]
  float* sb = sw + c1 * cin * 9;           // [c1]
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  for (int i = threadIdx.x; i < cin * HYW * HXW; i += NT) {
    const int ci = i / (HYW * HXW), r = (i / HXW) % HYW, c = i % HXW;
    const int yy = y0 + r - 1, xx = x0 + c - 1;
    sx[i] = (yy >= 0 && yy < Hx && xx >= 0 && xx < Wx) ? x[(((size_t)b * V + idx.v[ci]) * Hx + yy) * Wx + xx] : 0.f;
  }
  for (int i = threadIdx.x; i < c1 * cin * 9; i += NT) sw[i] = w1[i];
  for (int i = threadIdx.x; i < c1; i += NT) sb[i] = b1[i];
  __syncthreads();
  const int lx = threadIdx.x % TX, ly = threadIdx.x / TX;
  const int gx = x0 + lx, gy = y0 + ly;
  if (gx >= Wx || gy >= Hx) return;
  for (int oc0 = 0; oc0 < c1; oc0 += 16) {
    float acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = (oc0 + o < c1) ? sb[oc0 + o] : 0.f;
    for (int ci = 0; ci < cin; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float v = sx[(ci * HYW + ly + t / 3) * HXW + lx + t % 3];
#pragma unroll
        for (int o = 0; o < 16; ++o)
          if (oc0 + o < c1) acc[o] = fmaf(v, sw[(oc0 + o) * cin * 9 + ci * 9 + t], acc[o]);
      }
#pragma unroll
    for (int o = 0; o < 16; ++o)
      if (oc0 + o < c1) h1[(((size_t)b * c1 + oc0 + o) * Hx + gy) * Wx + gx] = from_f<T>(acc[o]);
  }
}

// g1 = PixelShuffle(mag)(GELU(h1)) for the generic path (the register-tiled conv1 kernel writes it from its epilogue)
template <typename T>
__global__ void gelu_shuffle_kernel(const T* __restrict__ h1, T* __restrict__ g1, int c1, int Hx, int Wx, int mag, size_t total) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int xl = (int)(i % Wx), yl = (int)((i / Wx) % Hx), ch = (int)((i / ((size_t)Wx * Hx)) % c1);
  const size_t b = i / ((size_t)Wx * Hx * c1);
  const int m2 = mag * mag, c4 = ch / m2, py = (ch % m2) / mag, px = ch % mag;
  g1[((b * (c1 / m2) + c4) * ((size_t)Hx * mag) + (size_t)yl * mag + py) * ((size_t)Wx * mag) + (size_t)xl * mag + px] =
      from_f<T>(gelu_f(to_f(h1[i])));
}

// dw1[oc][ci][tap] += sum dh1[b,oc,y,x] x7[b,ci,y+dy-1,x+dx-1]; db1[oc] += sum dh1.  Persistent CTAs, register partials.
template <typename T>
__global__ void __launch_bounds__(NT) conv1_bwd_kernel(const float* __restrict__ x, IdxList idx, const T* __restrict__ dh1,
                                                       float* __restrict__ dw1, float* __restrict__ db1, int B, int V,
                                                       int Hx, int Wx, int cin, int c1) {
  constexpr int PX = 16, PY = 8;           // 128-pixel tiles
  extern __shared__ float smem[];
  float* sx = smem;                        // [cin][PY+2][PX+2]
  float* sd = sx + cin * (PY + 2) * (PX + 2);   // [c1][PY*PX]
  const int ntap = cin * 9;                // <= 144
  // thread -> (oc group of 16, tap slot): tap slots per group = NT / (c1/16)
  const int ngrp = (c1 + 15) / 16;
  const int slots = NT / ngrp;             // 64 for c1=64
  const int grp = threadIdx.x / slots, slot = threadIdx.x % slots;
  const int nrep = (ntap + 1 + slots - 1) / slots;   // tap slots each thread walks (+1 = bias slot)
  float acc[3][16];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[r][o] = 0.f;
  const int tiles_x = (Wx + PX - 1) / PX, tiles_y = (Hx + PY - 1) / PY;
  const long long ntiles = (long long)B * tiles_x * tiles_y;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = (int)(tile / (tiles_x * tiles_y));
    const int ty = (int)((tile / tiles_x) % tiles_y), tx = (int)(tile % tiles_x);
    const int x0 = tx * PX, y0 = ty * PY;
    __syncthreads();
    for (int i = threadIdx.x; i < cin * (PY + 2) * (PX + 2); i += NT) {
      const int ci = i / ((PY + 2) * (PX + 2)), r = (i / (PX + 2)) % (PY + 2), c = i % (PX + 2);
      const int yy = y0 + r - 1, xx = x0 + c - 1;
      sx[i] = (yy >= 0 && yy < Hx && xx >= 0 && xx < Wx) ? x[(((size_t)b * V + idx.v[ci]) * Hx + yy) * Wx + xx] : 0.f;
    }
    for (int i = threadIdx.x; i < c1 * PY * PX; i += NT) {
      const int oc = i / (PY * PX), r = (i / PX) % PY, c = i % PX;
      const int yy = y0 + r, xx = x0 + c;
      sd[i] = (yy < Hx && xx < Wx) ? to_f(dh1[(((size_t)b * c1 + oc) * Hx + yy) * Wx + xx]) : 0.f;
    }
    __syncthreads();
    if (grp < ngrp) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if (r >= nrep) break;
        const int j = slot + r * slots;
        if (j > ntap) continue;
        const int ci = j / 9, t = j % 9;
        for (int pix = 0; pix < PY * PX; ++pix) {
          const int py = pix / PX, px = pix % PX;
          const float v = (j == ntap) ? 1.f : sx[(ci * (PY + 2) + py + t / 3) * (PX + 2) + px + t % 3];
#pragma unroll
          for (int o = 0; o < 16; ++o) {
            const int oc = grp * 16 + o;
            if (oc < c1) acc[r][o] = fmaf(v, sd[oc * PY * PX + pix], acc[r][o]);
          }
        }
      }
    }
  }
  if (grp < ngrp) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (r >= nrep) break;
      const int j = slot + r * slots;
      if (j > ntap) continue;
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        const int oc = grp * 16 + o;
        if (oc >= c1) continue;
        if (j == ntap) atomicAdd(&db1[oc], acc[r][o]);
        else atomicAdd(&dw1[(size_t)oc * ntap + j], acc[r][o]);
      }
    }
  }
}

// ------------------------------------------------------------------ head tail (high-res)

// flat index of unpatchified pixel (c, y, x) inside one sample of head_out
__device__ __forceinline__ size_t unpatch_idx(const HtArgs& a, int c, int y, int x) {
  const int p = a.p;
  const size_t cell = (size_t)(y / p) * (a.Wo / p) + (x / p);
  return cell * (p * p * a.C) + (size_t)(y % p) * (p * a.C) + (size_t)(x % p) * a.C + c;
}
__device__ __forceinline__ size_t shuffle_idx(const HtArgs& a, int c4, int y, int x) {
  const int m = a.mag;
  return (((size_t)c4 * m * m + (size_t)(y % m) * m + (x % m)) * a.Hx + y / m) * a.Wx + x / m;
}

template <typename T>
__global__ void __launch_bounds__(NT) headtail_fwd_kernel(const HtArgs a) {
  __shared__ float simg[MAXC][HYW][HXW];
  __shared__ float ssh[MAXC][HYW][HXW];
  __shared__ float sw[MAXC * MAXC * 9 * 2 + 2 * MAXC];
  const int C = a.C, cr = a.cr;
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const T* ho = reinterpret_cast<const T*>(a.head_out) + (size_t)b * a.Ho * a.Wo * C;
  const T* h1 = reinterpret_cast<const T*>(a.h1) + (size_t)b * cr * a.mag * a.mag * a.Hx * a.Wx;
  float* swo = sw; float* sw2 = sw + C * C * 9; float* sbo = sw2 + C * cr * 9; float* sb2 = sbo + C;
  for (int i = threadIdx.x; i < C * C * 9; i += NT) swo[i] = a.w_out[i];
  for (int i = threadIdx.x; i < C * cr * 9; i += NT) sw2[i] = a.w2[i];
  if (threadIdx.x < C) { sbo[threadIdx.x] = a.b_out[threadIdx.x]; sb2[threadIdx.x] = a.b2[threadIdx.x]; }
  for (int i = threadIdx.x; i < HYW * HXW; i += NT) {
    const int r = i / HXW, c = i % HXW;
    const int y = y0 + r - 1, x = x0 + c - 1;
    const bool in_img = (y >= 0 && y < a.Ho && x >= 0 && x < a.Wo);
    const bool in_sh = (y >= 0 && y < a.Hs && x >= 0 && x < a.Ws);
    for (int ch = 0; ch < C; ++ch) simg[ch][r][c] = in_img ? to_f(ho[unpatch_idx(a, ch, y, x)]) : 0.f;
    for (int ch = 0; ch < cr; ++ch) ssh[ch][r][c] = in_sh ? gelu_f(to_f(h1[shuffle_idx(a, ch, y, x)])) : 0.f;
  }
  __syncthreads();
  const int lx = threadIdx.x % TX, ly = threadIdx.x / TX;
  const int x = x0 + lx, y = y0 + ly;
  if (x >= a.Wo || y >= a.Ho) return;
  for (int c = 0; c < C; ++c) {
    float acc = sbo[c] + sb2[c];
    for (int ci = 0; ci < C; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc = fmaf(swo[(c * C + ci) * 9 + t], simg[ci][ly + t / 3][lx + t % 3], acc);
    for (int ci = 0; ci < cr; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc = fmaf(sw2[(c * cr + ci) * 9 + t], ssh[ci][ly + t / 3][lx + t % 3], acc);
    reinterpret_cast<T*>(a.preds)[(((size_t)b * C + c) * a.Ho + y) * a.Wo + x] = from_f<T>(acc);
  }
}

// backward over the Hs x Ws domain (>= Ho x Wo): d_head_out, dh1, and the four weight gradients
template <typename T>
__global__ void __launch_bounds__(NT) headtail_bwd_kernel(const HtArgs a) {
  __shared__ float sdp[MAXC][HYW][HXW];    // dpreds halo
  __shared__ float simg[MAXC][HYW][HXW];
  __shared__ float ssh[MAXC][HYW][HXW];
  __shared__ float sw[MAXC * MAXC * 9 * 2];
  const int C = a.C, cr = a.cr;
  float* swo = sw; float* sw2 = sw + C * C * 9;
  for (int i = threadIdx.x; i < C * C * 9; i += NT) swo[i] = a.w_out[i];
  for (int i = threadIdx.x; i < C * cr * 9; i += NT) sw2[i] = a.w2[i];
  const int n_wo = C * C * 9, n_w2 = C * cr * 9;
  // weight-gradient element owned by this thread (tap-thread scheme)
  const int j = threadIdx.x;
  float wacc = 0.f, wacc2 = 0.f;   // second accumulator when n_wo + n_w2 + C > NT
  const int tiles_x = (a.Ws + TX - 1) / TX, tiles_y = (a.Hs + TY - 1) / TY;
  const long long ntiles = (long long)a.B * tiles_x * tiles_y;
  const int lx = threadIdx.x % TX, ly = threadIdx.x / TX;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = (int)(tile / (tiles_x * tiles_y));
    const int x0 = (int)(tile % tiles_x) * TX, y0 = (int)((tile / tiles_x) % tiles_y) * TY;
    const T* ho = reinterpret_cast<const T*>(a.head_out) + (size_t)b * a.Ho * a.Wo * C;
    const T* h1 = reinterpret_cast<const T*>(a.h1) + (size_t)b * cr * a.mag * a.mag * a.Hx * a.Wx;
    const T* dp = reinterpret_cast<const T*>(a.dpreds) + (size_t)b * C * a.Ho * a.Wo;
    __syncthreads();
    for (int i = threadIdx.x; i < HYW * HXW; i += NT) {
      const int r = i / HXW, c = i % HXW;
      const int y = y0 + r - 1, x = x0 + c - 1;
      const bool in_img = (y >= 0 && y < a.Ho && x >= 0 && x < a.Wo);
      const bool in_sh = (y >= 0 && y < a.Hs && x >= 0 && x < a.Ws);
      for (int ch = 0; ch < C; ++ch) {
        sdp[ch][r][c] = in_img ? to_f(dp[((size_t)ch * a.Ho + y) * a.Wo + x]) : 0.f;
        simg[ch][r][c] = in_img ? to_f(ho[unpatch_idx(a, ch, y, x)]) : 0.f;
      }
      for (int ch = 0; ch < cr; ++ch) ssh[ch][r][c] = in_sh ? gelu_f(to_f(h1[shuffle_idx(a, ch, y, x)])) : 0.f;
    }
    __syncthreads();
    const int x = x0 + lx, y = y0 + ly;
    // data gradients (transposed convolutions in gather form)
    if (x < a.Wo && y < a.Ho) {
      T* dho = reinterpret_cast<T*>(a.d_head_out) + (size_t)b * a.Ho * a.Wo * C;
      for (int ci = 0; ci < C; ++ci) {
        float acc = 0.f;
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int t = 0; t < 9; ++t) acc = fmaf(swo[(c * C + ci) * 9 + t], sdp[c][ly + 2 - t / 3][lx + 2 - t % 3], acc);
        dho[unpatch_idx(a, ci, y, x)] = from_f<T>(acc);
      }
    }
    if (x < a.Ws && y < a.Hs) {
      T* dh = reinterpret_cast<T*>(a.dh1) + (size_t)b * cr * a.mag * a.mag * a.Hx * a.Wx;
      for (int ci = 0; ci < cr; ++ci) {
        float acc = 0.f;
        for (int c = 0; c < C; ++c)
#pragma unroll
          for (int t = 0; t < 9; ++t) acc = fmaf(sw2[(c * cr + ci) * 9 + t], sdp[c][ly + 2 - t / 3][lx + 2 - t % 3], acc);
        const size_t k = shuffle_idx(a, ci, y, x);
        dh[k] = from_f<T>(acc * dgelu_f(to_f(h1[k])));
      }
    }
    // weight gradients: one (c, ci, tap) per thread, loop over the tile's pixels
    for (int e = j; e < n_wo + n_w2 + C; e += NT) {
      float s = 0.f;
      if (e < n_wo + n_w2) {
        const bool second = e >= n_wo;
        const int ee = second ? e - n_wo : e;
        const int cin = second ? cr : C;
        const int c = ee / (cin * 9), ci = (ee / 9) % cin, t = ee % 9;
        const float (*src)[HYW][HXW] = second ? ssh : simg;
        for (int py = 0; py < TY; ++py)
#pragma unroll 8
          for (int px = 0; px < TX; ++px) s = fmaf(sdp[c][py + 1][px + 1], src[ci][py + t / 3][px + t % 3], s);
      } else {
        const int c = e - n_wo - n_w2;
        for (int py = 0; py < TY; ++py)
          for (int px = 0; px < TX; ++px) s += sdp[c][py + 1][px + 1];
      }
      if (e == j) wacc += s; else wacc2 += s;
    }
  }
  for (int e = j, r = 0; e < n_wo + n_w2 + C; e += NT, ++r) {
    const float s = r == 0 ? wacc : wacc2;
    if (e < n_wo) atomicAdd(&a.dw_out[e], s);
    else if (e < n_wo + n_w2) atomicAdd(&a.dw2[e - n_wo], s);
    else { atomicAdd(&a.db_out[e - n_wo - n_w2], s); atomicAdd(&a.db2[e - n_wo - n_w2], s); }
  }
}

// O2_HEADTAIL_GENERIC=1 keeps the generic kernels (any p / mag / cr) on shapes the register-tiled ones would take: tests
bool force_generic() {
  const char* e = getenv("O2_HEADTAIL_GENERIC");
  return e && e[0] == '1';
}

int fill_ht(HtArgs& a, int B, int C, int gh, int gw, int p, int mag, int cr, int Hx, int Wx) {
  O2_REQUIRE(B > 0 && C > 0 && C <= MAXC && cr > 0 && cr <= MAXC, "headtail: C=%d / cr=%d out of range (<=%d)", C, cr, MAXC);
  O2_REQUIRE(gh > 0 && gw > 0 && p > 0 && mag > 0, "headtail: bad dims");
  O2_REQUIRE(C * C * 9 + C * cr * 9 + C <= 2 * NT, "headtail: too many conv weights for the tap-thread scheme");
  memset(&a, 0, sizeof(a));
  a.B = B; a.C = C; a.gh = gh; a.gw = gw; a.p = p; a.mag = mag; a.cr = cr; a.Hx = Hx; a.Wx = Wx;
  a.Ho = gh * p * mag; a.Wo = gw * p * mag; a.Hs = Hx * mag; a.Ws = Wx * mag;
  O2_REQUIRE(a.Hs >= a.Ho && a.Ws >= a.Wo, "headtail: residual branch %dx%d smaller than the ViT output %dx%d", a.Hs, a.Ws,
             a.Ho, a.Wo);
  O2_REQUIRE(B <= 65535, "headtail: batch too large");
  return O2_OK;
}

}  // namespace

extern "C" int o2_path2_conv1_fwd(const float* x, const int* ch_idx_host, const float* w1, const float* b1, void* h1, void* g1,
                                  int dtype, int B, int V, int Hx, int Wx, int cin, int c1, int mag, void* stream) {
  O2_REQUIRE(x && ch_idx_host && w1 && b1 && h1, "conv1_fwd: null pointer");
  O2_REQUIRE(cin > 0 && cin <= MAXCIN && c1 > 0 && c1 <= MAXC1, "conv1_fwd: cin=%d / c1=%d out of range", cin, c1);
  O2_REQUIRE(!g1 || (mag > 0 && c1 % (mag * mag) == 0), "conv1_fwd: g1 needs mag > 0 and c1 %% mag^2 == 0 (c1=%d mag=%d)", c1, mag);
  IdxList idx;
  for (int i = 0; i < cin; ++i) {
    O2_REQUIRE(ch_idx_host[i] >= 0 && ch_idx_host[i] < V, "conv1_fwd: channel index %d out of range", ch_idx_host[i]);
    idx.v[i] = ch_idx_host[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (!force_generic()) {
    const int rc = conv1_fwd_fast(x, idx, w1, b1, h1, g1, dtype, B, V, Hx, Wx, cin, c1, mag, st);
    if (rc != kNotApplicable) return rc;
  }
  const size_t smem = sizeof(float) * ((size_t)cin * HYW * HXW + (size_t)c1 * cin * 9 + c1);
  dim3 grid((Wx + TX - 1) / TX, (Hx + TY - 1) / TY, B);
  if (dtype == O2_F32) {
    O2_CUDA(cudaFuncSetAttribute(conv1_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1_fwd_kernel<float><<<grid, NT, smem, st>>>(x, idx, w1, b1, (float*)h1, B, V, Hx, Wx, cin, c1);
  } else if (dtype == O2_BF16) {
    O2_CUDA(cudaFuncSetAttribute(conv1_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1_fwd_kernel<__nv_bfloat16><<<grid, NT, smem, st>>>(x, idx, w1, b1, (__nv_bfloat16*)h1, B, V, Hx, Wx, cin, c1);
  } else O2_FAIL(O2_ERR_ARG, "conv1_fwd: bad dtype %d", dtype);
  O2_LAUNCH_CHECK();
  if (g1) {
    const size_t total = (size_t)B * c1 * Hx * Wx;
    const unsigned gsz = (unsigned)((total + 255) / 256);
    if (dtype == O2_F32) gelu_shuffle_kernel<float><<<gsz, 256, 0, st>>>((const float*)h1, (float*)g1, c1, Hx, Wx, mag, total);
    else gelu_shuffle_kernel<__nv_bfloat16><<<gsz, 256, 0, st>>>((const __nv_bfloat16*)h1, (__nv_bfloat16*)g1, c1, Hx, Wx, mag, total);
    O2_LAUNCH_CHECK();
  }
  return O2_OK;
}

extern "C" int o2_path2_conv1_bwd(const float* x, const int* ch_idx_host, const void* dh1, float* dw1, float* db1,
                                  int dtype, int B, int V, int Hx, int Wx, int cin, int c1, void* stream) {
  O2_REQUIRE(x && ch_idx_host && dh1 && dw1 && db1, "conv1_bwd: null pointer");
  O2_REQUIRE(cin > 0 && cin <= MAXCIN && c1 > 0 && c1 <= MAXC1, "conv1_bwd: cin=%d / c1=%d out of range", cin, c1);
  const int ngrp = (c1 + 15) / 16, slots = NT / ngrp;
  O2_REQUIRE((cin * 9 + 1 + slots - 1) / slots <= 3, "conv1_bwd: cin*9=%d too large for the register tiling", cin * 9);
  IdxList idx;
  for (int i = 0; i < cin; ++i) {
    O2_REQUIRE(ch_idx_host[i] >= 0 && ch_idx_host[i] < V, "conv1_bwd: channel index out of range");
    idx.v[i] = ch_idx_host[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (!force_generic()) {
    const int rc = conv1_bwd_fast(x, idx, dh1, dw1, db1, dtype, B, V, Hx, Wx, cin, c1, st);
    if (rc != kNotApplicable) return rc;
  }
  const size_t smem = sizeof(float) * ((size_t)cin * 10 * 18 + (size_t)c1 * 128);
  const long long ntiles = (long long)B * ((Wx + 15) / 16) * ((Hx + 7) / 8);
  long long grid = (long long)o2_num_sms() * 2;
  if (grid > ntiles) grid = ntiles;
  if (dtype == O2_F32) {
    O2_CUDA(cudaFuncSetAttribute(conv1_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1_bwd_kernel<float><<<(unsigned)grid, NT, smem, st>>>(x, idx, (const float*)dh1, dw1, db1, B, V, Hx, Wx, cin, c1);
  } else if (dtype == O2_BF16) {
    O2_CUDA(cudaFuncSetAttribute(conv1_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1_bwd_kernel<__nv_bfloat16><<<(unsigned)grid, NT, smem, st>>>(x, idx, (const __nv_bfloat16*)dh1, dw1, db1, B, V, Hx, Wx, cin, c1);
  } else O2_FAIL(O2_ERR_ARG, "conv1_bwd: bad dtype %d", dtype);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

extern "C" int o2_headtail_fwd(const void* head_out, const void* h1, const void* g1, const float* w_out, const float* b_out,
                               const float* w2, const float* b2, void* preds, int dtype, int B, int C, int gh, int gw,
                               int p, int mag, int cr, int Hx, int Wx, void* stream) {
  HtArgs a;
  int rc = fill_ht(a, B, C, gh, gw, p, mag, cr, Hx, Wx);
  if (rc) return rc;
  O2_REQUIRE(head_out && h1 && w_out && b_out && w2 && b2 && preds, "headtail_fwd: null pointer");
  a.head_out = head_out; a.h1 = h1; a.g1 = g1; a.w_out = w_out; a.b_out = b_out; a.w2 = w2; a.b2 = b2; a.preds = preds;
  dim3 grid((a.Wo + TX - 1) / TX, (a.Ho + TY - 1) / TY, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (!force_generic()) {
    rc = headtail_fwd_fast(a, dtype, st);
    if (rc != kNotApplicable) return rc;
  }
  if (dtype == O2_F32) headtail_fwd_kernel<float><<<grid, NT, 0, st>>>(a);
  else if (dtype == O2_BF16) headtail_fwd_kernel<__nv_bfloat16><<<grid, NT, 0, st>>>(a);
  else O2_FAIL(O2_ERR_ARG, "headtail_fwd: bad dtype %d", dtype);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

extern "C" int o2_headtail_bwd(const void* dpreds, const void* head_out, const void* h1, const void* g1, const float* w_out,
                               const float* w2, void* d_head_out, void* dh1, float* dw_out, float* db_out, float* dw2,
                               float* db2, int dtype, int B, int C, int gh, int gw, int p, int mag, int cr, int Hx, int Wx,
                               void* stream) {
  HtArgs a;
  int rc = fill_ht(a, B, C, gh, gw, p, mag, cr, Hx, Wx);
  if (rc) return rc;
  O2_REQUIRE(dpreds && head_out && h1 && w_out && w2 && d_head_out && dh1 && dw_out && db_out && dw2 && db2,
             "headtail_bwd: null pointer");
  a.dpreds = dpreds; a.head_out = head_out; a.h1 = h1; a.g1 = g1; a.w_out = w_out; a.w2 = w2; a.d_head_out = d_head_out; a.dh1 = dh1;
  a.dw_out = dw_out; a.db_out = db_out; a.dw2 = dw2; a.db2 = db2;
  const long long ntiles = (long long)B * ((a.Ws + TX - 1) / TX) * ((a.Hs + TY - 1) / TY);
  long long grid = (long long)o2_num_sms() * 4;
  if (grid > ntiles) grid = ntiles;
  cudaStream_t st = (cudaStream_t)stream;
  if (!force_generic()) {
    rc = headtail_bwd_fast(a, dtype, st);
    if (rc != kNotApplicable) return rc;
  }
  if (dtype == O2_F32) headtail_bwd_kernel<float><<<(unsigned)grid, NT, 0, st>>>(a);
  else if (dtype == O2_BF16) headtail_bwd_kernel<__nv_bfloat16><<<(unsigned)grid, NT, 0, st>>>(a);
  else O2_FAIL(O2_ERR_ARG, "headtail_bwd: bad dtype %d", dtype);
  O2_LAUNCH_CHECK();
  return O2_OK;
}
