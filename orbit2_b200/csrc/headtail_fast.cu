// Register-tiled versions of the head-tail / residual-branch kernels for the shapes every reference config uses
// (patch_size 2, superres_mag 4, cnn_ratio 4: configs/interm_*.yaml:36-38).  Same arithmetic as the generic kernels in
// headtail.cu (reference: res_slimvit.py:107-112 path2, :167-179 unpatchify, :331 conv_out, :333-336 crop-add).
//
// What changed against the generic kernels (0.52 / 1.39 / 0.34 / 0.50 ms at 117M, B = 8: 3-5 % of the HBM roofline,
// bound by one LDS per FMA):
//   * a thread owns a 2 x 4 pixel block (head tail) or a 4-pixel strip x 16 output channels (conv1) in registers, pixel
//     windows arrive as LDS.128 + LDS.64 and filter taps as warp-uniform LDS.128, so ~15 LDS feed ~200 FFMA;
//   * the unpatchified image is staged by ROW PAIRS: rows (2k, 2k+1) of one tile are a contiguous run of the head
//     output (cells of p*p*C values), read and written back (d_head_out) with 8/16-byte vectors;
//   * weight gradients: every thread keeps a private set of taps in registers over a persistent tile loop and
//     contributes once at the end (one shuffle reduction + atomics per CTA, or plain atomics for conv1).
#include "headtail.cuh"

namespace o2ht {
namespace {

constexpr int FX = 64, FY = 32;            // output tile (head tail)
constexpr int FW = FX + 4, FH = FY + 2;    // halo tile: 66 columns padded to 68 (16-byte rows), 34 rows
constexpr int FNT = 256;
constexpr int XLR = FX / 4 + 2, YLR = FY / 4 + 2;   // low-res extent of the branch below one halo tile (mag 4)

template <int C> constexpr int kWP = (C * 9 + 3) & ~3;

template <typename T> __device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 t;
  t.x = pack_bf16x2(v[0], v[1]);
  t.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}

// 4 x 6 window (rows r0.., columns c0.. with c0 % 4 == 0) of one halo plane
__device__ __forceinline__ void window46(const float* plane, int r0, int c0, float (&v)[4][6]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float* s = plane + (r0 + r) * FW + c0;
    const float4 a = *reinterpret_cast<const float4*>(s);
    const float2 b = *reinterpret_cast<const float2*>(s + 4);
    v[r][0] = a.x; v[r][1] = a.y; v[r][2] = a.z; v[r][3] = a.w; v[r][4] = b.x; v[r][5] = b.y;
  }
}

// Staging.  Index decoding is hoisted out of the per-element path (a thread keeps a fixed column / vector slot and
// walks rows), global loads are 8/16-byte vectors issued in batches before any of them is consumed.  The first
// version decoded every element with div/mod chains and applied GELU per staged value: 230 M of the kernel's 266 M
// warp instructions were staging (ncu, profiles/r01_headtail_ncu.md).
template <typename T> __device__ __forceinline__ float gelu_t(float x);
template <> __device__ __forceinline__ float gelu_t<float>(float x) { return gelu_f(x); }
template <> __device__ __forceinline__ float gelu_t<__nv_bfloat16>(float x) { return gelu_fast(x); }
template <typename T> __device__ __forceinline__ float dgelu_t(float x);
template <> __device__ __forceinline__ float dgelu_t<float>(float x) { return dgelu_f(x); }
template <> __device__ __forceinline__ float dgelu_t<__nv_bfloat16>(float x) { return dgelu_fast(x); }   // |err| 1.5e-7 << bf16

// unpatchified image (SURVEY 8/a14 index map, p = 2): halo tile of C planes from the row-pair runs of one sample.
// Thread = one 4-element vector slot of the run (cell, r0 fixed), loop over the row pairs.
template <typename T, int C>
__device__ __forceinline__ void stage_img(float* simg, const T* __restrict__ ho, int x0, int y0, int Ho, int Wo) {
  constexpr int CS = 4 * C, NCELL = FX / 2 + 2, NPAIR = FY / 2 + 2, NVEC = NCELL * C;
  constexpr int NG = FNT / NVEC, U = 3;
  static_assert(NG >= 1, "run longer than the CTA");
  const int slot = threadIdx.x % NVEC, grp = threadIdx.x / NVEC;
  if (grp >= NG) return;
  const int celli = slot / C, r0 = (slot - celli * C) * 4;
  const int Wc = Wo >> 1, Hc = Ho >> 1;
  const int kp0 = (y0 >> 1) - 1, cx = (x0 >> 1) - 1 + celli;
  const bool cx_ok = cx >= 0 && cx < Wc;
  int off[4];          // smem offset of element e for row pair 0 (can be negative: row -1)
  int ppv[4];
  bool xok[4];         // its column lies inside the halo
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int r = r0 + e;
    const int pp = r / (2 * C), qq = (r / C) & 1, cc = r % C;
    const int xx = 2 * celli + qq - 1;
    ppv[e] = pp;
    xok[e] = (xx >= 0 && xx < FX + 2);
    off[e] = (cc * FH + pp - 1) * FW + xx;
  }
  const T* src = ho + (size_t)cx * CS + r0;
  for (int k0 = grp; k0 < NPAIR; k0 += U * NG) {
    float v[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kpi = k0 + u * NG, kp = kp0 + kpi;
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      if (kpi < NPAIR && cx_ok && kp >= 0 && kp < Hc) ld4<T>(src + (size_t)kp * Wc * CS, v[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kpi = k0 + u * NG;
      if (kpi >= NPAIR) continue;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int yy = 2 * kpi + ppv[e] - 1;
        if (xok[e] && yy >= 0 && yy < FH) simg[off[e] + 2 * kpi * FW] = v[u][e];
      }
    }
  }
}

// plain [planes][H][W] tensor with W % 4 == 0 and x0 % 4 == 0 (dpreds, the activated branch g1, the input fields):
// halo tile rows y0-1 .., columns x0-1 .. x0+COLS-2 from aligned 4-element vectors covering [x0-4, x0-4+4 NV)
template <typename T, int ROWS, int COLS, int PITCH, typename PlaneFn>
__device__ __forceinline__ void stage_rows4(float* dst, int nplanes, PlaneFn plane_ptr, int x0, int y0, int H, int W) {
  constexpr int NV = (COLS + 2) / 4 + 1, NG = FNT / NV, U = 4;   // vector j covers halo columns 4j-3 .. 4j
  const int j = threadIdx.x % NV, grp = threadIdx.x / NV;
  if (grp >= NG) return;
  const int xs = x0 - 4 + 4 * j;
  const bool x_ok = xs >= 0 && xs < W;
  const int col0 = 4 * j - 3;
  const int total = nplanes * ROWS;
  for (int r0 = grp; r0 < total; r0 += U * NG) {
    float v[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int rr = r0 + u * NG;
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      if (rr < total) {
        const int c = rr / ROWS, r = rr - c * ROWS;
        const int y = y0 + r - 1;
        if (x_ok && y >= 0 && y < H) ld4<T>(plane_ptr(c) + (size_t)y * W + xs, v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int rr = r0 + u * NG;
      if (rr >= total) continue;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (col0 + e >= 0 && col0 + e < COLS) dst[rr * PITCH + col0 + e] = v[u][e];
    }
  }
}

// same tile from scalar loads (any W): fallback for grids whose rows are not vector-aligned
template <typename T, int ROWS, int COLS, int PITCH, typename PlaneFn>
__device__ __forceinline__ void stage_planes(float* dst, int nplanes, PlaneFn plane_ptr, int x0, int y0, int H, int W) {
  constexpr int U = 8;
  const int total = nplanes * ROWS * COLS;
  for (int i0 = threadIdx.x; i0 < total; i0 += U * FNT) {
    float raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * FNT;
      raw[u] = 0.f;
      if (i < total) {
        const int c = i / (ROWS * COLS), r = (i / COLS) % ROWS, col = i % COLS;
        const int y = y0 + r - 1, x = x0 + col - 1;
        if (y >= 0 && y < H && x >= 0 && x < W) raw[u] = to_f(plane_ptr(c)[(size_t)y * W + x]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * FNT;
      if (i < total) {
        const int c = i / (ROWS * COLS), r = (i / COLS) % ROWS, col = i % COLS;
        dst[(c * ROWS + r) * PITCH + col] = raw[u];
      }
    }
  }
}

template <typename T, int ROWS, int COLS, int PITCH, typename PlaneFn>
__device__ __forceinline__ void stage_tile(float* dst, int nplanes, PlaneFn plane_ptr, int x0, int y0, int H, int W) {
  if ((W & 3) == 0) stage_rows4<T, ROWS, COLS, PITCH>(dst, nplanes, plane_ptr, x0, y0, H, W);
  else stage_planes<T, ROWS, COLS, PITCH>(dst, nplanes, plane_ptr, x0, y0, H, W);
}

// packed pairs (v[j], v[j+1]) of a 6-wide window row: the two FFMA2 lanes of adjacent output pixels.  The even pairs are
// register-adjacent already (LDS.128 / LDS.64 results); the odd ones are materialised exactly once (volatile: ptxas
// otherwise re-creates them with two MOVs in front of every use)
__device__ __forceinline__ uint64_t pack2_once(float lo, float hi) {
  uint64_t r;
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void pairs5(const float (&v)[6], uint64_t (&vp)[5]) {
  vp[0] = ptx::pack2(v[0], v[1]);
  vp[1] = pack2_once(v[1], v[2]);
  vp[2] = ptx::pack2(v[2], v[3]);
  vp[3] = pack2_once(v[3], v[4]);
  vp[4] = ptx::pack2(v[4], v[5]);
}

// ------------------------------------------------------------------ head tail forward
// arithmetic in packed fp32x2 (FFMA2: the two lanes are horizontally adjacent output pixels, the tap is duplicated in
// shared memory so one LDS.128 delivers two ready-made operand pairs): half the FMA-pipe slots of a scalar FFMA loop
template <typename T, int C, int CR>
__global__ void __launch_bounds__(FNT, 2) headtail_fwd_fast_kernel(const HtArgs a) {
  constexpr int WP = kWP<C>, CI = C + CR;
  extern __shared__ __align__(16) float smem_f[];
  float* tiles = smem_f;                                                   // [C + CR][FH][FW]: image planes, then branch planes
  uint64_t* sw = reinterpret_cast<uint64_t*>(tiles + CI * FH * FW);        // [CI][WP] (w, w): taps of plane ci, all outputs
  float* sb = reinterpret_cast<float*>(sw + CI * WP);                      // [C]
  const int b = blockIdx.z, x0 = blockIdx.x * FX, y0 = blockIdx.y * FY;
  const T* ho = reinterpret_cast<const T*>(a.head_out) + (size_t)b * a.Ho * a.Wo * C;
  const T* g1 = reinterpret_cast<const T*>(a.g1) + (size_t)b * CR * a.Hs * a.Ws;
  for (int i = threadIdx.x; i < CI * WP; i += FNT) {
    const int ci = i / WP, j = i % WP;
    float w = 0.f;
    if (j < C * 9) {
      const int c = j / 9, t = j % 9;
      w = ci < C ? a.w_out[(c * C + ci) * 9 + t] : a.w2[(c * CR + ci - C) * 9 + t];
    }
    sw[i] = ptx::pack2(w, w);
  }
  if (threadIdx.x < C) sb[threadIdx.x] = a.b_out[threadIdx.x] + a.b2[threadIdx.x];
  stage_img<T, C>(tiles, ho, x0, y0, a.Ho, a.Wo);
  const size_t gplane = (size_t)a.Hs * a.Ws;
  stage_tile<T, FH, FX + 2, FW>(tiles + C * FH * FW, CR, [&](int c) { return g1 + c * gplane; }, x0, y0, a.Hs, a.Ws);
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  uint64_t acc[C][2][2];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const uint64_t bb = ptx::pack2(sb[c], sb[c]);
#pragma unroll
    for (int r = 0; r < 2; ++r) acc[c][r][0] = acc[c][r][1] = bb;
  }
#pragma unroll
  for (int ci = 0; ci < CI; ++ci) {
    float v[4][6];
    window46(tiles + ci * FH * FW, 2 * ty, 4 * tx, v);
    uint64_t vp[4][5];
#pragma unroll
    for (int r = 0; r < 4; ++r) pairs5(v[r], vp[r]);
    uint64_t w[WP];
#pragma unroll
    for (int k = 0; k < WP; k += 2) {
      const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(sw + ci * WP + k);
      w[k] = t.x; w[k + 1] = t.y;
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            acc[c][r][0] = ptx::fma2(w[c * 9 + dy * 3 + dx], vp[r + dy][dx], acc[c][r][0]);
            acc[c][r][1] = ptx::fma2(w[c * 9 + dy * 3 + dx], vp[r + dy][dx + 2], acc[c][r][1]);
          }
  }
  const int x = x0 + 4 * tx;
  if (x >= a.Wo) return;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int y = y0 + 2 * ty + r;
    if (y >= a.Ho) continue;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float o[4];
      ptx::unpack2(acc[c][r][0], o[0], o[1]);
      ptx::unpack2(acc[c][r][1], o[2], o[3]);
      st4<T>(reinterpret_cast<T*>(a.preds) + (((size_t)b * C + c) * a.Ho + y) * a.Wo + x, o);
    }
  }
}

// ------------------------------------------------------------------ head tail backward
// tiles over the Hs x Ws domain (>= Ho x Wo); planes: dpreds [0,C), image [C,2C), branch [2C,2C+CR)
template <typename T, int C, int CR>
__global__ void __launch_bounds__(FNT, 2) headtail_bwd_fast_kernel(const HtArgs a) {
  constexpr int CI = C + CR, WT = 10;             // 9 taps + 1 pad: (w, w) pairs, 16-byte rows
  static_assert(CI * 32 <= FNT, "one warp per input plane for the weight gradients");
  extern __shared__ __align__(16) float smem_f[];
  float* tiles = smem_f;                                                      // [2C + CR][FH][FW]
  uint64_t* swT = reinterpret_cast<uint64_t*>(tiles + (2 * C + CR) * FH * FW);   // [C][CI][WT]: taps of (output c, input plane ci)
  for (int i = threadIdx.x; i < C * CI * WT; i += FNT) {
    const int t = i % WT, ci = (i / WT) % CI, c = i / (WT * CI);
    float w = 0.f;
    if (t < 9) w = ci < C ? a.w_out[(c * C + ci) * 9 + t] : a.w2[(c * CR + ci - C) * 9 + t];
    swT[i] = ptx::pack2(w, w);
  }
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int wci = threadIdx.x >> 5, wrow = threadIdx.x & 31;    // weight-gradient role: input plane, tile row
  float wacc[C][9], bacc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    bacc[c] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) wacc[c][t] = 0.f;
  }
  const int tiles_x = (a.Ws + FX - 1) / FX, tiles_y = (a.Hs + FY - 1) / FY;
  const long long ntiles = (long long)a.B * tiles_x * tiles_y;
  const int Wc = a.Wo >> 1;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = (int)(tile / (tiles_x * tiles_y));
    const int x0 = (int)(tile % tiles_x) * FX, y0 = (int)((tile / tiles_x) % tiles_y) * FY;
    const T* ho = reinterpret_cast<const T*>(a.head_out) + (size_t)b * a.Ho * a.Wo * C;
    const T* h1 = reinterpret_cast<const T*>(a.h1) + (size_t)b * CR * 16 * a.Hx * a.Wx;
    const T* dp = reinterpret_cast<const T*>(a.dpreds) + (size_t)b * C * a.Ho * a.Wo;
    const T* g1 = reinterpret_cast<const T*>(a.g1) + (size_t)b * CR * a.Hs * a.Ws;
    __syncthreads();
    const size_t plane = (size_t)a.Ho * a.Wo, gplane = (size_t)a.Hs * a.Ws;
    stage_tile<T, FH, FX + 2, FW>(tiles, C, [&](int c) { return dp + c * plane; }, x0, y0, a.Ho, a.Wo);
    stage_img<T, C>(tiles + C * FH * FW, ho, x0, y0, a.Ho, a.Wo);
    stage_tile<T, FH, FX + 2, FW>(tiles + 2 * C * FH * FW, CR, [&](int c) { return g1 + c * gplane; }, x0, y0, a.Hs, a.Ws);
    __syncthreads();
    // ---- data gradients of the thread's 2 x 4 block (transposed convolutions in gather form, FFMA2 over pixel pairs)
    {
      uint64_t d[CI][2][2];
#pragma unroll
      for (int ci = 0; ci < CI; ++ci)
#pragma unroll
        for (int r = 0; r < 2; ++r) d[ci][r][0] = d[ci][r][1] = 0ull;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float v[4][6];
        window46(tiles + c * FH * FW, 2 * ty, 4 * tx, v);
        uint64_t vp[4][5];
#pragma unroll
        for (int r = 0; r < 4; ++r) pairs5(v[r], vp[r]);
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
          uint64_t w[WT];
#pragma unroll
          for (int k = 0; k < WT; k += 2) {
            const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(swT + (c * CI + ci) * WT + k);
            w[k] = t.x; w[k + 1] = t.y;
          }
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                d[ci][r][0] = ptx::fma2(w[dy * 3 + dx], vp[r + 2 - dy][2 - dx], d[ci][r][0]);
                d[ci][r][1] = ptx::fma2(w[dy * 3 + dx], vp[r + 2 - dy][4 - dx], d[ci][r][1]);
              }
        }
      }
      const int x = x0 + 4 * tx, y = y0 + 2 * ty;
      if (x < a.Wo && y < a.Ho) {
        // rows (y, y+1) x columns [x, x+4) = two whole cells of the head output: 8C contiguous values
        T* dho = reinterpret_cast<T*>(a.d_head_out) + (size_t)b * a.Ho * a.Wo * C + ((size_t)(y >> 1) * Wc + (x >> 1)) * (4 * C);
        float o[8 * C];
#pragma unroll
        for (int cell = 0; cell < 2; ++cell)
#pragma unroll
          for (int pp = 0; pp < 2; ++pp)
#pragma unroll
            for (int cc = 0; cc < C; ++cc)
              ptx::unpack2(d[cc][pp][cell], o[cell * 4 * C + pp * 2 * C + cc], o[cell * 4 * C + pp * 2 * C + C + cc]);
#pragma unroll
        for (int k = 0; k < 8 * C; k += 4) {
          const float q[4] = {o[k], o[k + 1], o[k + 2], o[k + 3]};
          st4<T>(dho + k, q);
        }
      }
      if (x < a.Ws) {
        T* dh = reinterpret_cast<T*>(a.dh1) + (size_t)b * CR * 16 * a.Hx * a.Wx;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int yy = y + r;
          if (yy >= a.Hs) continue;
#pragma unroll
          for (int c4 = 0; c4 < CR; ++c4) {
            float g[4];
            ptx::unpack2(d[C + c4][r][0], g[0], g[1]);
            ptx::unpack2(d[C + c4][r][1], g[2], g[3]);
            const size_t k0 = ((size_t)(c4 * 16 + (yy & 3) * 4) * a.Hx + (yy >> 2)) * a.Wx + (x >> 2);
            const size_t cs = (size_t)a.Hx * a.Wx;
            float hv[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) hv[p] = to_f(h1[k0 + p * cs]);
#pragma unroll
            for (int p = 0; p < 4; ++p) dh[k0 + p * cs] = from_f<T>(g[p] * dgelu_t<T>(hv[p]));
          }
        }
      }
    }
    // ---- weight gradients: warp wci owns input plane wci, lane = tile row, private 9 x C taps
    if (wci < CI) {
      const float* plane_w = tiles + (C + wci) * FH * FW;
#pragma unroll 1
      for (int s = 0; s < FX / 4; ++s) {
        float P[3][6], D[C][6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float* q = plane_w + (wrow + r) * FW + 4 * s;
          const float4 u = *reinterpret_cast<const float4*>(q);
          const float2 w = *reinterpret_cast<const float2*>(q + 4);
          P[r][0] = u.x; P[r][1] = u.y; P[r][2] = u.z; P[r][3] = u.w; P[r][4] = w.x; P[r][5] = w.y;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float* q = tiles + (c * FH + wrow + 1) * FW + 4 * s;
          const float4 u = *reinterpret_cast<const float4*>(q);
          const float2 w = *reinterpret_cast<const float2*>(q + 4);
          D[c][0] = u.x; D[c][1] = u.y; D[c][2] = u.z; D[c][3] = u.w; D[c][4] = w.x; D[c][5] = w.y;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int j = 0; j < 4; ++j) wacc[c][dy * 3 + dx] = fmaf(D[c][j + 1], P[dy][j + dx], wacc[c][dy * 3 + dx]);
          bacc[c] += (D[c][1] + D[c][2]) + (D[c][3] + D[c][4]);
        }
      }
    }
  }
  if (wci < CI) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float s = warp_sum(wacc[c][t]);
        if (wrow == 0) {
          if (wci < C) atomicAdd(&a.dw_out[(c * C + wci) * 9 + t], s);
          else atomicAdd(&a.dw2[(c * CR + wci - C) * 9 + t], s);
        }
      }
      if (wci == 0) {
        const float s = warp_sum(bacc[c]);
        if (wrow == 0) { atomicAdd(&a.db_out[c], s); atomicAdd(&a.db2[c], s); }
      }
    }
  }
}

template <int C, int CR> constexpr size_t fwd_smem() { return sizeof(float) * ((C + CR) * FH * FW + 2 * (C + CR) * kWP<C> + C + 4); }
template <int C, int CR> constexpr size_t bwd_smem() { return sizeof(float) * ((2 * C + CR) * FH * FW + 2 * C * (C + CR) * 10); }

template <typename T, int C, int CR>
int launch_fwd(const HtArgs& a, cudaStream_t st) {
  constexpr size_t smem = fwd_smem<C, CR>();
  O2_CUDA(cudaFuncSetAttribute(headtail_fwd_fast_kernel<T, C, CR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((a.Wo + FX - 1) / FX, (a.Ho + FY - 1) / FY, a.B);
  headtail_fwd_fast_kernel<T, C, CR><<<grid, FNT, smem, st>>>(a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}
template <typename T, int C, int CR>
int launch_bwd(const HtArgs& a, cudaStream_t st) {
  constexpr size_t smem = bwd_smem<C, CR>();
  O2_CUDA(cudaFuncSetAttribute(headtail_bwd_fast_kernel<T, C, CR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntiles = (long long)a.B * ((a.Ws + FX - 1) / FX) * ((a.Hs + FY - 1) / FY);
  long long grid = (long long)o2_num_sms() * 2;
  if (grid > ntiles) grid = ntiles;
  headtail_bwd_fast_kernel<T, C, CR><<<(unsigned)grid, FNT, smem, st>>>(a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

bool fast_shape(const HtArgs& a) { return a.g1 && a.p == 2 && a.mag == 4 && a.cr == 4 && a.C >= 1 && a.C <= 4; }

// ------------------------------------------------------------------ path2 conv1 (low-res grid), c1 = 64
constexpr int QX = 64, QY = 16, QW = QX + 4;

// forward: thread = 4-pixel strip x 16 output channels at a time; the two FFMA2 lanes are two OUTPUT CHANNELS of one
// pixel, so the taps [ci][dy][dx][oc] come straight out of warp-uniform LDS.128 as operand pairs and only the six
// window values need a (v, v) duplicate
template <typename T>
__global__ void __launch_bounds__(FNT, 2) conv1_fwd_fast_kernel(const float* __restrict__ x, IdxList idx, const float* __restrict__ w1,
                                                               const float* __restrict__ b1, T* __restrict__ h1, T* __restrict__ g1,
                                                               int B, int V, int Hx, int Wx, int cin, int c1) {
  extern __shared__ __align__(16) float smem_f[];
  float* sx = smem_f;                              // [cin][QY + 2][QW]
  float* sw = sx + cin * (QY + 2) * QW;            // [cin][3][3][c1]
  float* sb = sw + cin * 9 * c1;                   // [c1]
  const int b = blockIdx.z, x0 = blockIdx.x * QX, y0 = blockIdx.y * QY;
  const size_t plane = (size_t)Hx * Wx;
  const float* xb = x + (size_t)b * V * plane;
  stage_tile<float, QY + 2, QX + 2, QW>(sx, cin, [&](int c) { return xb + idx.v[c] * plane; }, x0, y0, Hx, Wx);
  for (int i = threadIdx.x; i < cin * 9 * c1; i += FNT) {      // w1 [oc][ci][dy][dx] -> sw [ci][dy][dx][oc]
    const int oc = i % c1, t = (i / c1) % 9, ci = i / (9 * c1);
    sw[i] = w1[((size_t)oc * cin + ci) * 9 + t];
  }
  for (int i = threadIdx.x; i < c1; i += FNT) sb[i] = b1[i];
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int gy = y0 + ty, gx = x0 + 4 * tx;
  if (gy >= Hx || gx >= Wx) return;
  const bool vec = ((Wx & 3) == 0);                // then gx + 3 < Wx and the strip is 8/16-byte aligned
  for (int oc0 = 0; oc0 < c1; oc0 += 16) {
    uint64_t acc[8][4];                            // [channel pair][pixel]
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[o][p] = ptx::pack2(sb[oc0 + 2 * o], sb[oc0 + 2 * o + 1]);
    for (int ci = 0; ci < cin; ++ci) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const float* s = sx + (ci * (QY + 2) + ty + dy) * QW + 4 * tx;
        const float4 u = *reinterpret_cast<const float4*>(s);
        const float2 w = *reinterpret_cast<const float2*>(s + 4);
        const uint64_t vv[6] = {pack2_once(u.x, u.x), pack2_once(u.y, u.y), pack2_once(u.z, u.z),
                                pack2_once(u.w, u.w), pack2_once(w.x, w.x), pack2_once(w.y, w.y)};
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float* wp = sw + ((ci * 3 + dy) * 3 + dx) * c1 + oc0;      // 16 consecutive output channels
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const ulonglong2 f = *reinterpret_cast<const ulonglong2*>(wp + 4 * q);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              acc[2 * q][p] = ptx::fma2(f.x, vv[p + dx], acc[2 * q][p]);
              acc[2 * q + 1][p] = ptx::fma2(f.y, vv[p + dx], acc[2 * q + 1][p]);
            }
          }
        }
      }
    }
    float r[16][4];                                // [channel in the group = py * 4 + px][pixel]
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
      for (int p = 0; p < 4; ++p) ptx::unpack2(acc[o][p], r[2 * o][p], r[2 * o + 1][p]);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      T* dst = h1 + (((size_t)b * c1 + oc0 + k) * Hx + gy) * Wx + gx;
      if (vec) {
        st4<T>(dst, r[k]);
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (gx + p < Wx) dst[p] = from_f<T>(r[k][p]);
      }
    }
    if (g1) {
      // PixelShuffle(4): the 16 channels of this group are the 4 x 4 sub-pixels of plane oc0 / 16, so the strip's
      // activations are 4 rows of 16 contiguous high-res values
      const int Ws = Wx * 4;
      T* gb = g1 + (((size_t)b * (c1 >> 4) + (oc0 >> 4)) * (Hx * 4) + 4 * gy) * Ws + 4 * gx;
#pragma unroll
      for (int py = 0; py < 4; ++py) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          float q[4];
#pragma unroll
          for (int px = 0; px < 4; ++px) q[px] = gelu_t<T>(r[py * 4 + px][p]);
          if (vec) {
            st4<T>(gb + (size_t)py * Ws + 4 * p, q);
          } else if (gx + p < Wx) {
#pragma unroll
            for (int px = 0; px < 4; ++px) gb[(size_t)py * Ws + 4 * p + px] = from_f<T>(q[px]);
          }
        }
      }
    }
  }
}

// backward (weight gradients): warp = 8 output channels, lane = (input plane ci, tap row dy) -> 8 x 3 private taps kept
// over a persistent loop of 64 x 4 pixel tiles (FFMA2 over pixel pairs, the two halves are added at the end); nobody
// shares a tap, so the CTA ends with plain atomics.  The bias gradient rides along for free: lane cin*3 runs the same
// instruction stream on a row of ones, so its first tap accumulates sum(dh1).
constexpr int RY = 4;
template <typename T>
__global__ void __launch_bounds__(FNT, 2) conv1_bwd_fast_kernel(const float* __restrict__ x, IdxList idx, const T* __restrict__ dh1,
                                                               float* __restrict__ dw1, float* __restrict__ db1, int B, int V,
                                                               int Hx, int Wx, int cin, int c1) {
  extern __shared__ __align__(16) float smem_f[];
  float* sx = smem_f;                              // [cin][RY + 2][QW], then one row of ones
  float* sones = sx + cin * (RY + 2) * QW;         // [QW]
  float* sd = sones + QW;                          // [c1][RY][QX]
  const int og = threadIdx.x >> 5, cd = threadIdx.x & 31;   // warp = group of 8 output channels, lane = ci * 3 + dy
  const int ci = cd / 3, dy = cd % 3;
  const bool bias_lane = (cd == cin * 3);
  const bool active = (og < (c1 >> 3)) && (cd <= cin * 3);
  for (int i = threadIdx.x; i < QW; i += FNT) sones[i] = 1.f;
  uint64_t acc[8][3];
#pragma unroll
  for (int o = 0; o < 8; ++o)
#pragma unroll
    for (int d = 0; d < 3; ++d) acc[o][d] = 0ull;
  const int tiles_x = (Wx + QX - 1) / QX, tiles_y = (Hx + RY - 1) / RY;
  const long long ntiles = (long long)B * tiles_x * tiles_y;
  const size_t plane = (size_t)Hx * Wx;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = (int)(tile / (tiles_x * tiles_y));
    const int x0 = (int)(tile % tiles_x) * QX, y0 = (int)((tile / tiles_x) % tiles_y) * RY;
    const float* xb = x + (size_t)b * V * plane;
    const T* db = dh1 + (size_t)b * c1 * plane;
    __syncthreads();
    stage_tile<float, RY + 2, QX + 2, QW>(sx, cin, [&](int c) { return xb + idx.v[c] * plane; }, x0, y0, Hx, Wx);
    if ((Wx & 3) == 0) {
      // dh1 tile (no halo) through aligned 4-element vectors: [c1][RY][QX / 4]
      constexpr int U = 8;
      const int total = c1 * RY * (QX / 4);
      for (int i0 = threadIdx.x; i0 < total; i0 += U * FNT) {
        float v[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * FNT;
          v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
          const int j = i % (QX / 4), r = (i / (QX / 4)) % RY, oc = i / (RY * (QX / 4));
          const int yy = y0 + r, xx = x0 + 4 * j;
          if (i < total && yy < Hx && xx < Wx) ld4<T>(db + (size_t)oc * plane + (size_t)yy * Wx + xx, v[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * FNT;
          if (i < total) *reinterpret_cast<float4*>(sd + 4 * i) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
        }
      }
    } else {
      // rows y0 .., columns x0 .. (stage_planes offsets by -1, hence the +1 origin)
      stage_planes<T, RY, QX, QX>(sd, c1, [&](int c) { return db + c * plane; }, x0 + 1, y0 + 1, Hx, Wx);
    }
    __syncthreads();
    if (active) {
#pragma unroll 1
      for (int r = 0; r < RY; ++r) {
        const float* srow = bias_lane ? sones : sx + (ci * (RY + 2) + r + dy) * QW;
        const float* drow = sd + ((og * 8) * RY + r) * QX;
#pragma unroll 2
        for (int s = 0; s < QX / 4; ++s) {
          const float4 u = *reinterpret_cast<const float4*>(srow + 4 * s);
          const float2 w = *reinterpret_cast<const float2*>(srow + 4 * s + 4);
          const float v[6] = {u.x, u.y, u.z, u.w, w.x, w.y};
          uint64_t vp[5];
          pairs5(v, vp);
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const ulonglong2 g = *reinterpret_cast<const ulonglong2*>(drow + o * RY * QX + 4 * s);   // (g0, g1), (g2, g3)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              acc[o][d] = ptx::fma2(g.x, vp[d], acc[o][d]);
              acc[o][d] = ptx::fma2(g.y, vp[d + 2], acc[o][d]);
            }
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int oc = og * 8 + o;
      if (bias_lane) {
        float lo, hi;
        ptx::unpack2(acc[o][0], lo, hi);
        atomicAdd(&db1[oc], lo + hi);
        continue;
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float lo, hi;
        ptx::unpack2(acc[o][d], lo, hi);
        atomicAdd(&dw1[((size_t)oc * cin + ci) * 9 + dy * 3 + d], lo + hi);
      }
    }
  }
}

}  // namespace

#define O2_HT_DISPATCH(fn, T)                                                  \
  switch (a.C) {                                                              \
    case 1: return fn<T, 1, 4>(a, st);                                        \
    case 2: return fn<T, 2, 4>(a, st);                                        \
    case 3: return fn<T, 3, 4>(a, st);                                        \
    default: return fn<T, 4, 4>(a, st);                                       \
  }

int headtail_fwd_fast(const HtArgs& a, int dtype, cudaStream_t st) {
  if (!fast_shape(a)) return kNotApplicable;
  if (dtype == O2_F32) { O2_HT_DISPATCH(launch_fwd, float) }
  if (dtype == O2_BF16) { O2_HT_DISPATCH(launch_fwd, __nv_bfloat16) }
  return kNotApplicable;
}

int headtail_bwd_fast(const HtArgs& a, int dtype, cudaStream_t st) {
  if (!fast_shape(a)) return kNotApplicable;
  if (dtype == O2_F32) { O2_HT_DISPATCH(launch_bwd, float) }
  if (dtype == O2_BF16) { O2_HT_DISPATCH(launch_bwd, __nv_bfloat16) }
  return kNotApplicable;
}

int conv1_fwd_fast(const float* x, const IdxList& idx, const float* w1, const float* b1, void* h1, void* g1, int dtype, int B,
                   int V, int Hx, int Wx, int cin, int c1, int mag, cudaStream_t st) {
  if ((c1 & 15) != 0 || B > 65535 || (g1 && mag != 4)) return kNotApplicable;
  const size_t smem = sizeof(float) * ((size_t)cin * (QY + 2) * QW + (size_t)cin * 9 * c1 + c1);
  if (smem > 100 * 1024) return kNotApplicable;
  dim3 grid((Wx + QX - 1) / QX, (Hx + QY - 1) / QY, B);
  if (dtype == O2_F32) {
    O2_CUDA(cudaFuncSetAttribute(conv1_fwd_fast_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1_fwd_fast_kernel<float><<<grid, FNT, smem, st>>>(x, idx, w1, b1, (float*)h1, (float*)g1, B, V, Hx, Wx, cin, c1);
  } else if (dtype == O2_BF16) {
    O2_CUDA(cudaFuncSetAttribute(conv1_fwd_fast_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1_fwd_fast_kernel<__nv_bfloat16><<<grid, FNT, smem, st>>>(x, idx, w1, b1, (__nv_bfloat16*)h1, (__nv_bfloat16*)g1, B, V, Hx, Wx, cin, c1);
  } else return kNotApplicable;
  O2_LAUNCH_CHECK();
  return O2_OK;
}

int conv1_bwd_fast(const float* x, const IdxList& idx, const void* dh1, float* dw1, float* db1, int dtype, int B, int V,
                   int Hx, int Wx, int cin, int c1, cudaStream_t st) {
  if ((c1 & 7) != 0 || (c1 >> 3) > FNT / 32 || cin * 3 > 31) return kNotApplicable;
  const size_t smem = sizeof(float) * ((size_t)cin * (RY + 2) * QW + QW + (size_t)c1 * RY * QX);
  if (smem > 100 * 1024) return kNotApplicable;
  const long long ntiles = (long long)B * ((Wx + QX - 1) / QX) * ((Hx + RY - 1) / RY);
  long long grid = (long long)o2_num_sms() * 2;
  if (grid > ntiles) grid = ntiles;
  if (dtype == O2_F32) {
    O2_CUDA(cudaFuncSetAttribute(conv1_bwd_fast_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1_bwd_fast_kernel<float><<<(unsigned)grid, FNT, smem, st>>>(x, idx, (const float*)dh1, dw1, db1, B, V, Hx, Wx, cin, c1);
  } else if (dtype == O2_BF16) {
    O2_CUDA(cudaFuncSetAttribute(conv1_bwd_fast_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv1_bwd_fast_kernel<__nv_bfloat16><<<(unsigned)grid, FNT, smem, st>>>(x, idx, (const __nv_bfloat16*)dh1, dw1, db1, B, V, Hx, Wx, cin, c1);
  } else return kNotApplicable;
  O2_LAUNCH_CHECK();
  return O2_OK;
}

}  // namespace o2ht
