// fp32 SIMT flash attention (forward, backward) -- the exact-parity arm (fp32 <= 1e-4) of o2_attn_*.
// reference: components/attention.py:50-78 (softmax(q k^T * hd^-0.5) v, bidirectional, no mask).
// qkv [B,N,3,heads,hd], out [B,N,heads,hd], lse [B,heads,N] (natural log).
//
// 64 x 64 (query x key) tiles staged in shared memory, 256 threads each owning a 4 x 4 sub-block of S and a
// 4 x (HD/16) sub-block of the output; online softmax with half-warp shuffles.  Backward = preprocess (delta),
// one kernel that owns a KV tile (dK, dV) and one that owns a Q tile (dQ): no atomics, deterministic.
#include "common.cuh"

namespace {

constexpr int BQ = 64, BKV = 64, NT = 256;
// row padding of the shared-memory tiles (bank-conflict-free transposing stores); head dim 128 only fits the 227 KiB of
// shared memory unpadded (backward: 4 transposed + 2 row-major operand tiles), trading store conflicts for capacity
template <int HD> constexpr int kPad = (HD <= 64) ? 4 : 0;

struct AttnArgs {
  const float* qkv; float* out; float* lse;
  const float* dout; float* dqkv; float* delta;
  int B, N, heads, hd; float scale;
  ptx::AttnDrop drop;        // thr16 == 0: no attention dropout
};

// scaled keep factor of element (q, k) of (batch, head) bh: 0 or 1 / keep_prob (mask function: common.cuh)
__device__ __forceinline__ float drop_factor(const AttnArgs& a, uint32_t key_bh, int q, int k) {
  const uint32_t w = ptx::attn_keep_word(a.drop, key_bh, (uint32_t)q, (uint32_t)k >> 5);
  return ((w >> ptx::attn_keep_bit((uint32_t)k)) & 1u) ? a.drop.inv_keep : 0.f;
}

__device__ __forceinline__ const float* qkv_ptr(const AttnArgs& a, int b, int which, int h) {
  return a.qkv + ((size_t)b * a.N * 3 + which) * a.heads * a.hd + (size_t)h * a.hd;
}
// row stride of q/k/v inside qkv
__device__ __forceinline__ size_t qkv_rs(const AttnArgs& a) { return (size_t)3 * a.heads * a.hd; }

// Unpadded transposed tiles of the wide-head kernels (pitch 64 floats = 0 mod 32 banks, no room for padding at head dim
// 256): a thread that stored dst[d][row] element-wise with consecutive d would hit ONE bank 32 times.  Instead every
// thread moves a 4 x 4 block -- four 16-byte global loads along d, transposed in registers, four 16-byte shared stores --
// and the 4-row group g of column d lives at group g ^ ((d >> 2) & 15): the 32 threads of a warp (consecutive d groups)
// then cover all banks (4 wavefronts of 128 bytes, the minimum for 512 bytes).  Readers XOR their row-group index with tswz.
template <int HD> constexpr bool kSwz = (HD >= 256);
__device__ __forceinline__ int tswz(int d) { return (d >> 2) & 15; }
template <int W>     // W = number of columns (head-dim elements) of the chunk; dst is [W][64] floats, unpadded
__device__ __forceinline__ void load_T_swizzled(float* dst, const float* src, size_t row_stride, int row0, int N) {
  static_assert(W % 4 == 0, "4-column groups");
  for (int it = threadIdx.x; it < 16 * (W / 4); it += NT) {
    const int dq = it % (W / 4), rq = it / (W / 4);
    const int d0 = dq * 4, r0 = rq * 4;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int gr = row0 + r0 + k;
      v[k] = (gr < N) ? *reinterpret_cast<const float4*>(src + (size_t)gr * row_stride + d0) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* base = dst + (size_t)d0 * 64 + ((rq ^ (dq & 15)) << 2);          // tswz(d0 + c) == dq & 15 for c = 0..3
    *reinterpret_cast<float4*>(base) = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
    *reinterpret_cast<float4*>(base + 64) = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
    *reinterpret_cast<float4*>(base + 128) = make_float4(v[0].z, v[1].z, v[2].z, v[3].z);
    *reinterpret_cast<float4*>(base + 192) = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
  }
}

// load a [64 rows x HD] tile transposed into smem: dst[d][row] (row pitch 64+pad), rows >= n_valid zero
template <int HD>
__device__ __forceinline__ void load_tile_T(float (*dst)[BQ + kPad<HD>], const float* src, size_t row_stride, int row0, int N) {
  if constexpr (kSwz<HD>) {
    static_assert(kPad<HD> == 0, "the swizzled layout assumes a 64-float pitch");
    load_T_swizzled<HD>(&dst[0][0], src, row_stride, row0, N);
  } else {
    for (int i = threadIdx.x; i < 64 * HD; i += NT) {
      const int r = i / HD, d = i % HD;
      const int gr = row0 + r;
      dst[d][r] = (gr < N) ? src[(size_t)gr * row_stride + d] : 0.f;
    }
  }
}
template <int HD>
__device__ __forceinline__ void load_tile(float (*dst)[HD + kPad<HD>], const float* src, size_t row_stride, int row0, int N) {
  static_assert(HD % 4 == 0 && kPad<HD> % 4 == 0, "16-byte rows");
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (row_stride & 3) == 0) {       // 16-byte vectors (block-uniform branch)
    for (int i = threadIdx.x; i < 64 * (HD / 4); i += NT) {
      const int r = i / (HD / 4), d = (i % (HD / 4)) * 4;
      const int gr = row0 + r;
      *reinterpret_cast<float4*>(&dst[r][d]) =
          (gr < N) ? *reinterpret_cast<const float4*>(src + (size_t)gr * row_stride + d) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  for (int i = threadIdx.x; i < 64 * HD; i += NT) {
    const int r = i / HD, d = i % HD;
    const int gr = row0 + r;
    dst[r][d] = (gr < N) ? src[(size_t)gr * row_stride + d] : 0.f;
  }
}

template <int HD>
__global__ void __launch_bounds__(NT) attn_fwd_kernel(const AttnArgs a) {
  constexpr int DC = HD / 16;
  extern __shared__ float smem[];
  float (*qT)[BQ + kPad<HD>] = reinterpret_cast<float (*)[BQ + kPad<HD>]>(smem);                       // [HD][68]
  float (*kT)[BKV + kPad<HD>] = reinterpret_cast<float (*)[BKV + kPad<HD>]>(smem + HD * (BQ + kPad<HD>));     // [HD][68]
  float (*vS)[HD + kPad<HD>] = reinterpret_cast<float (*)[HD + kPad<HD>]>(smem + 2 * HD * (BQ + kPad<HD>));   // [64][HD+4]
  float (*pT)[BQ + kPad<HD>] = reinterpret_cast<float (*)[BQ + kPad<HD>]>(smem + 2 * HD * (BQ + kPad<HD>) + BKV * (HD + kPad<HD>));  // [64 key][68]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int b = blockIdx.y / a.heads, h = blockIdx.y % a.heads;
  const int q0 = blockIdx.x * BQ;
  const size_t rs = qkv_rs(a);
  load_tile_T<HD>(qT, qkv_ptr(a, b, 0, h), rs, q0, a.N);
  float o[4][DC] = {};
  float m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; }
  for (int k0 = 0; k0 < a.N; k0 += BKV) {
    __syncthreads();
    load_tile_T<HD>(kT, qkv_ptr(a, b, 1, h), rs, k0, a.N);
    load_tile<HD>(vS, qkv_ptr(a, b, 2, h), rs, k0, a.N);
    __syncthreads();
    float s[4][4] = {};
#pragma unroll 8
    for (int d = 0; d < HD; ++d) {
      const int sw = kSwz<HD> ? tswz(d) : 0;               // wide-head tiles are stored swizzled (load_T_swizzled)
      const float4 qv = *reinterpret_cast<const float4*>(&qT[d][(ty ^ sw) * 4]);
      const float4 kv = *reinterpret_cast<const float4*>(&kT[d][(tx ^ sw) * 4]);
      const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + tx * 4 + j < a.N) ? s[i][j] * a.scale : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mn = fmaxf(m[i], mx);
      const float alpha = __expf(m[i] - mn);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = __expf(s[i][j] - mn); sum += s[i][j]; }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      l[i] = l[i] * alpha + sum;
      m[i] = mn;
#pragma unroll
      for (int c = 0; c < DC; ++c) o[i][c] *= alpha;
      if (a.drop.thr16 > 0) {                  // P V uses the masked probabilities, l the unmasked ones
        const uint32_t key_bh = ptx::attn_drop_key(a.drop, blockIdx.y);
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] *= drop_factor(a, key_bh, q0 + ty * 4 + i, k0 + tx * 4 + j);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) pT[tx * 4 + j][ty * 4 + i] = s[i][j];
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < BKV; ++k) {
      const float4 pv = *reinterpret_cast<const float4*>(&pT[k][ty * 4]);
      const float pa[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        const float vv = vS[k][c * 16 + tx];            // thread tx owns head-dim columns c * 16 + tx: 16 banks, no conflict
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i][c] = fmaf(pa[i], vv, o[i][c]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
    if (row >= a.N) continue;
    const float inv = 1.f / l[i];
    float* op = a.out + ((size_t)b * a.N + row) * a.heads * a.hd + (size_t)h * a.hd + tx;
#pragma unroll
    for (int c = 0; c < DC; ++c) op[c * 16] = o[i][c] * inv;
    if (tx == 0) a.lse[((size_t)b * a.heads + h) * a.N + row] = m[i] + __logf(l[i]);
  }
}

// delta[b,h,n] = sum_d dout * out
__global__ void attn_delta_kernel(const float* __restrict__ out, const float* __restrict__ dout, float* __restrict__ delta,
                                  int B, int N, int heads, int hd) {
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * N * heads;
  if (warp >= total) return;
  const int h = (int)(warp % heads);
  const long long bn = warp / heads;
  const int n = (int)(bn % N);
  const int b = (int)(bn / N);
  const float* o = out + warp * hd;
  const float* d = dout + warp * hd;
  float s = 0.f;
  for (int i = lane; i < hd; i += 32) s += o[i] * d[i];
  s = warp_sum(s);
  if (lane == 0) delta[((size_t)b * heads + h) * N + n] = s;
}

// Shared S / dS recomputation for the two backward kernels.  On return: p[i][j] and ds[i][j] for rows ty*4+i, keys tx*4+j.
template <int HD>
__device__ __forceinline__ void recompute_p_ds(const AttnArgs& a, float (*qT)[BQ + kPad<HD>], float (*kT)[BKV + kPad<HD>],
                                               float (*doT)[BQ + kPad<HD>], float (*vT)[BKV + kPad<HD>], const float* lse_s,
                                               const float* delta_s, int q0, int k0, int tx, int ty, float (&p)[4][4],
                                               float (&ds)[4][4]) {
  float s[4][4] = {}, dp[4][4] = {};
#pragma unroll 8
  for (int d = 0; d < HD; ++d) {
    const float4 qv = *reinterpret_cast<const float4*>(&qT[d][ty * 4]);
    const float4 kv = *reinterpret_cast<const float4*>(&kT[d][tx * 4]);
    const float4 gv = *reinterpret_cast<const float4*>(&doT[d][ty * 4]);
    const float4 vv = *reinterpret_cast<const float4*>(&vT[d][tx * 4]);
    const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
    const float ga[4] = {gv.x, gv.y, gv.z, gv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
        dp[i][j] = fmaf(ga[i], va[j], dp[i][j]);
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = (q0 + ty * 4 + i < a.N) && (k0 + tx * 4 + j < a.N);
      p[i][j] = ok ? __expf(s[i][j] * a.scale - lse_s[ty * 4 + i]) : 0.f;
      if (a.drop.thr16 > 0) {                  // dP = dP_drop o M; the returned p is the masked one (it feeds dV = P_drop^T dO)
        const float f = drop_factor(a, ptx::attn_drop_key(a.drop, blockIdx.y), q0 + ty * 4 + i, k0 + tx * 4 + j);
        ds[i][j] = p[i][j] * (dp[i][j] * f - delta_s[ty * 4 + i]) * a.scale;
        p[i][j] *= f;
      } else {
        ds[i][j] = p[i][j] * (dp[i][j] - delta_s[ty * 4 + i]) * a.scale;
      }
    }
}

// owns one KV tile: dK, dV
template <int HD>
__global__ void __launch_bounds__(NT) attn_bwd_kv_kernel(const AttnArgs a) {
  constexpr int DC = HD / 16;
  extern __shared__ float smem[];
  float (*qT)[BQ + kPad<HD>] = reinterpret_cast<float (*)[BQ + kPad<HD>]>(smem);
  float (*kT)[BKV + kPad<HD>] = qT + HD;
  float (*doT)[BQ + kPad<HD>] = kT + HD;
  float (*vT)[BKV + kPad<HD>] = doT + HD;
  float* after = smem + 4 * HD * (BQ + kPad<HD>);
  float (*qS)[HD + kPad<HD>] = reinterpret_cast<float (*)[HD + kPad<HD>]>(after);              // [64 row][HD+4]
  float (*doS)[HD + kPad<HD>] = qS + BQ;
  float (*pS)[BKV + kPad<HD>] = reinterpret_cast<float (*)[BKV + kPad<HD>]>(after + 2 * BQ * (HD + kPad<HD>));   // [row][key]
  float (*dsS)[BKV + kPad<HD>] = pS + BQ;
  float* lse_s = reinterpret_cast<float*>(dsS + BQ);
  float* delta_s = lse_s + BQ;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int b = blockIdx.y / a.heads, h = blockIdx.y % a.heads;
  const int k0 = blockIdx.x * BKV;
  const size_t rs = qkv_rs(a);
  const size_t os = (size_t)a.heads * a.hd;
  load_tile_T<HD>(kT, qkv_ptr(a, b, 1, h), rs, k0, a.N);
  load_tile_T<HD>(vT, qkv_ptr(a, b, 2, h), rs, k0, a.N);
  float dk[4][DC] = {}, dv[4][DC] = {};   // keys ty*4+i, dims tx*DC+c
  for (int q0 = 0; q0 < a.N; q0 += BQ) {
    __syncthreads();
    load_tile_T<HD>(qT, qkv_ptr(a, b, 0, h), rs, q0, a.N);
    load_tile<HD>(qS, qkv_ptr(a, b, 0, h), rs, q0, a.N);
    const float* dob = a.dout + (size_t)b * a.N * os + (size_t)h * a.hd;
    load_tile_T<HD>(doT, dob, os, q0, a.N);
    load_tile<HD>(doS, dob, os, q0, a.N);
    if (threadIdx.x < BQ) {
      const int r = q0 + threadIdx.x;
      lse_s[threadIdx.x] = r < a.N ? a.lse[((size_t)b * a.heads + h) * a.N + r] : 0.f;
      delta_s[threadIdx.x] = r < a.N ? a.delta[((size_t)b * a.heads + h) * a.N + r] : 0.f;
    }
    __syncthreads();
    float p[4][4], ds[4][4];
    recompute_p_ds<HD>(a, qT, kT, doT, vT, lse_s, delta_s, q0, k0, tx, ty, p, ds);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { pS[ty * 4 + i][tx * 4 + j] = p[i][j]; dsS[ty * 4 + i][tx * 4 + j] = ds[i][j]; }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < BQ; ++r) {
      const float4 pv = *reinterpret_cast<const float4*>(&pS[r][ty * 4]);
      const float4 sv = *reinterpret_cast<const float4*>(&dsS[r][ty * 4]);
      const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, sa[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        const float g = doS[r][tx * DC + c], qv = qS[r][tx * DC + c];
#pragma unroll
        for (int i = 0; i < 4; ++i) { dv[i][c] = fmaf(pa[i], g, dv[i][c]); dk[i][c] = fmaf(sa[i], qv, dk[i][c]); }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int key = k0 + ty * 4 + i;
    if (key >= a.N) continue;
    float* dkp = a.dqkv + (((size_t)b * a.N + key) * 3 + 1) * os + (size_t)h * a.hd + tx * DC;
    float* dvp = a.dqkv + (((size_t)b * a.N + key) * 3 + 2) * os + (size_t)h * a.hd + tx * DC;
#pragma unroll
    for (int c = 0; c < DC; ++c) { dkp[c] = dk[i][c]; dvp[c] = dv[i][c]; }
  }
}

// owns one Q tile: dQ
template <int HD>
__global__ void __launch_bounds__(NT) attn_bwd_q_kernel(const AttnArgs a) {
  constexpr int DC = HD / 16;
  extern __shared__ float smem[];
  float (*qT)[BQ + kPad<HD>] = reinterpret_cast<float (*)[BQ + kPad<HD>]>(smem);
  float (*kT)[BKV + kPad<HD>] = qT + HD;
  float (*doT)[BQ + kPad<HD>] = kT + HD;
  float (*vT)[BKV + kPad<HD>] = doT + HD;
  float* after = smem + 4 * HD * (BQ + kPad<HD>);
  float (*kS)[HD + kPad<HD>] = reinterpret_cast<float (*)[HD + kPad<HD>]>(after);                       // [64 key][HD+4]
  float (*dsT)[BQ + kPad<HD>] = reinterpret_cast<float (*)[BQ + kPad<HD>]>(after + BKV * (HD + kPad<HD>));     // [key][row]
  float* lse_s = reinterpret_cast<float*>(dsT + BKV);
  float* delta_s = lse_s + BQ;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int b = blockIdx.y / a.heads, h = blockIdx.y % a.heads;
  const int q0 = blockIdx.x * BQ;
  const size_t rs = qkv_rs(a);
  const size_t os = (size_t)a.heads * a.hd;
  load_tile_T<HD>(qT, qkv_ptr(a, b, 0, h), rs, q0, a.N);
  load_tile_T<HD>(doT, a.dout + (size_t)b * a.N * os + (size_t)h * a.hd, os, q0, a.N);
  if (threadIdx.x < BQ) {
    const int r = q0 + threadIdx.x;
    lse_s[threadIdx.x] = r < a.N ? a.lse[((size_t)b * a.heads + h) * a.N + r] : 0.f;
    delta_s[threadIdx.x] = r < a.N ? a.delta[((size_t)b * a.heads + h) * a.N + r] : 0.f;
  }
  float dq[4][DC] = {};
  for (int k0 = 0; k0 < a.N; k0 += BKV) {
    __syncthreads();
    load_tile_T<HD>(kT, qkv_ptr(a, b, 1, h), rs, k0, a.N);
    load_tile<HD>(kS, qkv_ptr(a, b, 1, h), rs, k0, a.N);
    load_tile_T<HD>(vT, qkv_ptr(a, b, 2, h), rs, k0, a.N);
    __syncthreads();
    float p[4][4], ds[4][4];
    recompute_p_ds<HD>(a, qT, kT, doT, vT, lse_s, delta_s, q0, k0, tx, ty, p, ds);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dsT[tx * 4 + j][ty * 4 + i] = ds[i][j];
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < BKV; ++k) {
      const float4 sv = *reinterpret_cast<const float4*>(&dsT[k][ty * 4]);
      const float sa[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        const float kv = kS[k][tx * DC + c];
#pragma unroll
        for (int i = 0; i < 4; ++i) dq[i][c] = fmaf(sa[i], kv, dq[i][c]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
    if (row >= a.N) continue;
    float* dqp = a.dqkv + (((size_t)b * a.N + row) * 3 + 0) * os + (size_t)h * a.hd + tx * DC;
#pragma unroll
    for (int c = 0; c < DC; ++c) dqp[c] = dq[i][c];
  }
}

// ---- wide heads (interm_10b: head dim 256).  The four transposed operand tiles of the S / dP recomputation would need
// 256 KiB, so they are staged in chunks of HC head-dim columns and S / dP accumulate over the chunks; the row-major
// tiles for the gradient contractions and the register accumulators keep the full head dim.
template <int HC>
__device__ __forceinline__ void accumulate_s_dp(float (*qT)[BQ], float (*kT)[BKV], float (*doT)[BQ], float (*vT)[BKV], int tx,
                                                int ty, float (&s)[4][4], float (&dp)[4][4]) {
#pragma unroll 8
  for (int d = 0; d < HC; ++d) {
    const int sw = tswz(d);                                 // chunk tiles are stored swizzled (load_T_swizzled)
    const float4 qv = *reinterpret_cast<const float4*>(&qT[d][(ty ^ sw) * 4]);
    const float4 kv = *reinterpret_cast<const float4*>(&kT[d][(tx ^ sw) * 4]);
    const float4 gv = *reinterpret_cast<const float4*>(&doT[d][(ty ^ sw) * 4]);
    const float4 vv = *reinterpret_cast<const float4*>(&vT[d][(tx ^ sw) * 4]);
    const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
    const float ga[4] = {gv.x, gv.y, gv.z, gv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
        dp[i][j] = fmaf(ga[i], va[j], dp[i][j]);
      }
  }
}
__device__ __forceinline__ void finish_p_ds(const AttnArgs& a, const float* lse_s, const float* delta_s, int q0, int k0, int tx,
                                            int ty, const float (&s)[4][4], const float (&dp)[4][4], float (&p)[4][4],
                                            float (&ds)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = (q0 + ty * 4 + i < a.N) && (k0 + tx * 4 + j < a.N);
      p[i][j] = ok ? __expf(s[i][j] * a.scale - lse_s[ty * 4 + i]) : 0.f;
      if (a.drop.thr16 > 0) {
        const float f = drop_factor(a, ptx::attn_drop_key(a.drop, blockIdx.y), q0 + ty * 4 + i, k0 + tx * 4 + j);
        ds[i][j] = p[i][j] * (dp[i][j] * f - delta_s[ty * 4 + i]) * a.scale;
        p[i][j] *= f;
      } else {
        ds[i][j] = p[i][j] * (dp[i][j] - delta_s[ty * 4 + i]) * a.scale;
      }
    }
}
// [64 rows x HC] chunk (columns c0 ..) transposed into dst[d][row] (no padding, swizzled row groups: load_T_swizzled)
template <int HC>
__device__ __forceinline__ void load_chunk_T(float (*dst)[BQ], const float* src, size_t row_stride, int row0, int N) {
  load_T_swizzled<HC>(&dst[0][0], src, row_stride, row0, N);
}

template <int HD, int HC>
__global__ void __launch_bounds__(NT) attn_bwd_kv_wide_kernel(const AttnArgs a) {
  constexpr int DC = HD / 16;
  extern __shared__ float smem[];
  float (*qT)[BQ] = reinterpret_cast<float (*)[BQ]>(smem);      // chunk buffers [HC][64]
  float (*kT)[BKV] = qT + HC;
  float (*doT)[BQ] = kT + HC;
  float (*vT)[BKV] = doT + HC;
  float* after = smem + 4 * HC * BQ;
  float (*qS)[HD] = reinterpret_cast<float (*)[HD]>(after);      // [64 row][HD]
  float (*doS)[HD] = qS + BQ;
  float (*pS)[BKV] = reinterpret_cast<float (*)[BKV]>(after + 2 * BQ * HD);   // [row][key]
  float (*dsS)[BKV] = pS + BQ;
  float* lse_s = reinterpret_cast<float*>(dsS + BQ);
  float* delta_s = lse_s + BQ;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int b = blockIdx.y / a.heads, h = blockIdx.y % a.heads;
  const int k0 = blockIdx.x * BKV;
  const size_t rs = qkv_rs(a);
  const size_t os = (size_t)a.heads * a.hd;
  float dk[4][DC] = {}, dv[4][DC] = {};   // keys ty*4+i, dims tx*DC+c
  const float* dob = a.dout + (size_t)b * a.N * os + (size_t)h * a.hd;
  for (int q0 = 0; q0 < a.N; q0 += BQ) {
    __syncthreads();
    load_tile<HD>(reinterpret_cast<float (*)[HD + kPad<HD>]>(qS), qkv_ptr(a, b, 0, h), rs, q0, a.N);
    load_tile<HD>(reinterpret_cast<float (*)[HD + kPad<HD>]>(doS), dob, os, q0, a.N);
    if (threadIdx.x < BQ) {
      const int r = q0 + threadIdx.x;
      lse_s[threadIdx.x] = r < a.N ? a.lse[((size_t)b * a.heads + h) * a.N + r] : 0.f;
      delta_s[threadIdx.x] = r < a.N ? a.delta[((size_t)b * a.heads + h) * a.N + r] : 0.f;
    }
    float s[4][4] = {}, dp[4][4] = {};
    for (int c0 = 0; c0 < HD; c0 += HC) {
      __syncthreads();
      load_chunk_T<HC>(qT, qkv_ptr(a, b, 0, h) + c0, rs, q0, a.N);
      load_chunk_T<HC>(kT, qkv_ptr(a, b, 1, h) + c0, rs, k0, a.N);
      load_chunk_T<HC>(doT, dob + c0, os, q0, a.N);
      load_chunk_T<HC>(vT, qkv_ptr(a, b, 2, h) + c0, rs, k0, a.N);
      __syncthreads();
      accumulate_s_dp<HC>(qT, kT, doT, vT, tx, ty, s, dp);
    }
    float p[4][4], ds[4][4];
    finish_p_ds(a, lse_s, delta_s, q0, k0, tx, ty, s, dp, p, ds);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { pS[ty * 4 + i][tx * 4 + j] = p[i][j]; dsS[ty * 4 + i][tx * 4 + j] = ds[i][j]; }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < BQ; ++r) {
      const float4 pv = *reinterpret_cast<const float4*>(&pS[r][ty * 4]);
      const float4 sv = *reinterpret_cast<const float4*>(&dsS[r][ty * 4]);
      const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, sa[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        const float g = doS[r][c * 16 + tx], qv = qS[r][c * 16 + tx];   // columns c * 16 + tx: conflict-free
#pragma unroll
        for (int i = 0; i < 4; ++i) { dv[i][c] = fmaf(pa[i], g, dv[i][c]); dk[i][c] = fmaf(sa[i], qv, dk[i][c]); }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int key = k0 + ty * 4 + i;
    if (key >= a.N) continue;
    float* dkp = a.dqkv + (((size_t)b * a.N + key) * 3 + 1) * os + (size_t)h * a.hd + tx;
    float* dvp = a.dqkv + (((size_t)b * a.N + key) * 3 + 2) * os + (size_t)h * a.hd + tx;
#pragma unroll
    for (int c = 0; c < DC; ++c) { dkp[c * 16] = dk[i][c]; dvp[c * 16] = dv[i][c]; }
  }
}

template <int HD, int HC>
__global__ void __launch_bounds__(NT) attn_bwd_q_wide_kernel(const AttnArgs a) {
  constexpr int DC = HD / 16;
  extern __shared__ float smem[];
  float (*qT)[BQ] = reinterpret_cast<float (*)[BQ]>(smem);
  float (*kT)[BKV] = qT + HC;
  float (*doT)[BQ] = kT + HC;
  float (*vT)[BKV] = doT + HC;
  float* after = smem + 4 * HC * BQ;
  float (*kS)[HD] = reinterpret_cast<float (*)[HD]>(after);                       // [64 key][HD]
  float (*dsT)[BQ] = reinterpret_cast<float (*)[BQ]>(after + BKV * HD);           // [key][row]
  float* lse_s = reinterpret_cast<float*>(dsT + BKV);
  float* delta_s = lse_s + BQ;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int b = blockIdx.y / a.heads, h = blockIdx.y % a.heads;
  const int q0 = blockIdx.x * BQ;
  const size_t rs = qkv_rs(a);
  const size_t os = (size_t)a.heads * a.hd;
  const float* dob = a.dout + (size_t)b * a.N * os + (size_t)h * a.hd;
  if (threadIdx.x < BQ) {
    const int r = q0 + threadIdx.x;
    lse_s[threadIdx.x] = r < a.N ? a.lse[((size_t)b * a.heads + h) * a.N + r] : 0.f;
    delta_s[threadIdx.x] = r < a.N ? a.delta[((size_t)b * a.heads + h) * a.N + r] : 0.f;
  }
  float dq[4][DC] = {};
  for (int k0 = 0; k0 < a.N; k0 += BKV) {
    __syncthreads();
    load_tile<HD>(reinterpret_cast<float (*)[HD + kPad<HD>]>(kS), qkv_ptr(a, b, 1, h), rs, k0, a.N);
    float s[4][4] = {}, dp[4][4] = {};
    for (int c0 = 0; c0 < HD; c0 += HC) {
      __syncthreads();
      load_chunk_T<HC>(qT, qkv_ptr(a, b, 0, h) + c0, rs, q0, a.N);
      load_chunk_T<HC>(kT, qkv_ptr(a, b, 1, h) + c0, rs, k0, a.N);
      load_chunk_T<HC>(doT, dob + c0, os, q0, a.N);
      load_chunk_T<HC>(vT, qkv_ptr(a, b, 2, h) + c0, rs, k0, a.N);
      __syncthreads();
      accumulate_s_dp<HC>(qT, kT, doT, vT, tx, ty, s, dp);
    }
    float p[4][4], ds[4][4];
    finish_p_ds(a, lse_s, delta_s, q0, k0, tx, ty, s, dp, p, ds);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dsT[tx * 4 + j][ty * 4 + i] = ds[i][j];
    __syncthreads();
#pragma unroll 2
    for (int k = 0; k < BKV; ++k) {
      const float4 sv = *reinterpret_cast<const float4*>(&dsT[k][ty * 4]);
      const float sa[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        const float kv = kS[k][c * 16 + tx];
#pragma unroll
        for (int i = 0; i < 4; ++i) dq[i][c] = fmaf(sa[i], kv, dq[i][c]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
    if (row >= a.N) continue;
    float* dqp = a.dqkv + (((size_t)b * a.N + row) * 3 + 0) * os + (size_t)h * a.hd + tx;
#pragma unroll
    for (int c = 0; c < DC; ++c) dqp[c * 16] = dq[i][c];
  }
}

template <int HD, int HC> int run_bwd_wide(const AttnArgs& a, cudaStream_t st) {
  constexpr size_t kv_smem = sizeof(float) * (4 * HC * BQ + 2 * BQ * HD + 2 * BQ * BKV + 2 * BQ);
  constexpr size_t q_smem = sizeof(float) * (4 * HC * BQ + BKV * HD + BKV * BQ + 2 * BQ);
  static_assert(kv_smem <= 227 * 1024 && kPad<HD> == 0, "wide-head tiles must fit the 227 KiB of shared memory");
  O2_CUDA(cudaFuncSetAttribute(attn_bwd_kv_wide_kernel<HD, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kv_smem));
  O2_CUDA(cudaFuncSetAttribute(attn_bwd_q_wide_kernel<HD, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q_smem));
  const long long warps = (long long)a.B * a.N * a.heads;
  attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(a.out, a.dout, a.delta, a.B, a.N, a.heads, a.hd);
  O2_LAUNCH_CHECK();
  attn_bwd_kv_wide_kernel<HD, HC><<<dim3((a.N + BKV - 1) / BKV, a.B * a.heads), NT, kv_smem, st>>>(a);
  O2_LAUNCH_CHECK();
  attn_bwd_q_wide_kernel<HD, HC><<<dim3((a.N + BQ - 1) / BQ, a.B * a.heads), NT, q_smem, st>>>(a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

template <int HD> constexpr size_t fwd_smem() { return sizeof(float) * (2 * HD * (BQ + kPad<HD>) + BKV * (HD + kPad<HD>) + BKV * (BQ + kPad<HD>)); }
template <int HD> constexpr size_t bwd_kv_smem() {
  return sizeof(float) * (4 * HD * (BQ + kPad<HD>) + 2 * BQ * (HD + kPad<HD>) + 2 * BQ * (BKV + kPad<HD>) + 2 * BQ);
}
template <int HD> constexpr size_t bwd_q_smem() {
  return sizeof(float) * (4 * HD * (BQ + kPad<HD>) + BKV * (HD + kPad<HD>) + BKV * (BQ + kPad<HD>) + 2 * BQ);
}

template <int HD> int run_fwd(const AttnArgs& a, cudaStream_t st) {
  O2_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<HD>()));
  attn_fwd_kernel<HD><<<dim3((a.N + BQ - 1) / BQ, a.B * a.heads), NT, fwd_smem<HD>(), st>>>(a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}
template <int HD> int run_bwd(const AttnArgs& a, cudaStream_t st) {
  O2_CUDA(cudaFuncSetAttribute(attn_bwd_kv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_kv_smem<HD>()));
  O2_CUDA(cudaFuncSetAttribute(attn_bwd_q_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_q_smem<HD>()));
  const long long warps = (long long)a.B * a.N * a.heads;
  attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(a.out, a.dout, a.delta, a.B, a.N, a.heads, a.hd);
  O2_LAUNCH_CHECK();
  attn_bwd_kv_kernel<HD><<<dim3((a.N + BKV - 1) / BKV, a.B * a.heads), NT, bwd_kv_smem<HD>(), st>>>(a);
  O2_LAUNCH_CHECK();
  attn_bwd_q_kernel<HD><<<dim3((a.N + BQ - 1) / BQ, a.B * a.heads), NT, bwd_q_smem<HD>(), st>>>(a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

}  // namespace

ptx::AttnDrop make_drop_simt(float p, uint64_t seed, uint32_t site, int N) { return ptx::make_attn_drop(p, seed, site, N); }

int o2_attn_fwd_simt(const void* qkv, void* out, float* lse, int B, int N, int heads, int hd, float scale, float p_drop,
                     uint64_t seed, uint32_t site, cudaStream_t st) {
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.drop = make_drop_simt(p_drop, seed, site, N);
  a.qkv = (const float*)qkv; a.out = (float*)out; a.lse = lse; a.B = B; a.N = N; a.heads = heads; a.hd = hd; a.scale = scale;
  O2_REQUIRE((long long)B * heads <= 65535, "attn: B*heads too large");
  if (hd == 32) return run_fwd<32>(a, st);
  if (hd == 64) return run_fwd<64>(a, st);
  if (hd == 128) return run_fwd<128>(a, st);
  if (hd == 256) {                                    // interm_10b: 32 heads x 256 (208 KiB of tiles, no padding)
    O2_REQUIRE(((uintptr_t)qkv & 15) == 0, "attn_simt: head dim 256 needs a 16-byte aligned qkv (vector tile loads)");
    return run_fwd<256>(a, st);
  }
  O2_FAIL(O2_ERR_UNSUPPORTED, "attn_simt: head dim %d not in {32,64,128,256}", hd);
}

int o2_attn_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta, int B,
                     int N, int heads, int hd, float scale, float p_drop, uint64_t seed, uint32_t site, cudaStream_t st) {
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.drop = make_drop_simt(p_drop, seed, site, N);
  a.qkv = (const float*)qkv; a.out = (float*)const_cast<void*>(out); a.lse = const_cast<float*>(lse);
  a.dout = (const float*)dout; a.dqkv = (float*)dqkv; a.delta = delta;
  a.B = B; a.N = N; a.heads = heads; a.hd = hd; a.scale = scale;
  O2_REQUIRE((long long)B * heads <= 65535, "attn: B*heads too large");
  if (hd == 32) return run_bwd<32>(a, st);
  if (hd == 64) return run_bwd<64>(a, st);
  if (hd == 128) return run_bwd<128>(a, st);
  if (hd == 256) {
    O2_REQUIRE((((uintptr_t)qkv | (uintptr_t)dout) & 15) == 0, "attn_simt: head dim 256 needs 16-byte aligned qkv / dout");
    return run_bwd_wide<256, 64>(a, st);
  }
  O2_FAIL(O2_ERR_UNSUPPORTED, "attn_simt backward: head dim %d not in {32,64,128,256}", hd);
}

int o2_attn_fwd_tc(const void* qkv, void* out, float* lse, int B, int N, int heads, int hd, float scale, float p_drop,
                   uint64_t seed, uint32_t site, cudaStream_t st);
int o2_attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta, int B,
                   int N, int heads, int hd, float scale, int parts, float p_drop, uint64_t seed, uint32_t site,
                   cudaStream_t st);

extern "C" int o2_attn_fwd_drop(int impl, const void* qkv, void* out, float* lse, int B, int N, int heads, int hd, float scale,
                                float p_drop, uint64_t seed, uint32_t site, void* stream) {
  O2_REQUIRE(qkv && out && lse, "attn_fwd: null pointer");
  O2_REQUIRE(B > 0 && N > 0 && heads > 0 && hd > 0, "attn_fwd: bad dims");
  O2_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attn_fwd: p_drop=%f outside [0, 1)", (double)p_drop);
  if (impl == O2_GEMM_SIMT_F32)
    return o2_attn_fwd_simt(qkv, out, lse, B, N, heads, hd, scale, p_drop, seed, site, (cudaStream_t)stream);
  if (impl == O2_GEMM_TC_BF16)
    return o2_attn_fwd_tc(qkv, out, lse, B, N, heads, hd, scale, p_drop, seed, site, (cudaStream_t)stream);
  O2_FAIL(O2_ERR_ARG, "attn_fwd: unknown impl %d", impl);
}

extern "C" int o2_attn_fwd(int impl, const void* qkv, void* out, float* lse, int B, int N, int heads, int hd, float scale,
                           void* stream) {
  return o2_attn_fwd_drop(impl, qkv, out, lse, B, N, heads, hd, scale, 0.f, 0, 0, stream);
}

extern "C" int o2_attn_bwd_parts_drop(int impl, int parts, const void* qkv, const void* out, const void* dout,
                                      const float* lse, void* dqkv, float* delta, int B, int N, int heads, int hd,
                                      float scale, float p_drop, uint64_t seed, uint32_t site, void* stream) {
  O2_REQUIRE(qkv && out && dout && lse && dqkv && delta, "attn_bwd: null pointer");
  O2_REQUIRE(B > 0 && N > 0 && heads > 0 && hd > 0, "attn_bwd: bad dims");
  O2_REQUIRE(parts > 0 && parts <= O2_ATTN_BWD_ALL, "attn_bwd: bad parts mask %d", parts);
  O2_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attn_bwd: p_drop=%f outside [0, 1)", (double)p_drop);
  if (impl == O2_GEMM_SIMT_F32) {
    O2_REQUIRE(parts == O2_ATTN_BWD_ALL, "attn_bwd: the fp32 SIMT path runs all parts in one call");
    return o2_attn_bwd_simt(qkv, out, dout, lse, dqkv, delta, B, N, heads, hd, scale, p_drop, seed, site, (cudaStream_t)stream);
  }
  if (impl == O2_GEMM_TC_BF16)
    return o2_attn_bwd_tc(qkv, out, dout, lse, dqkv, delta, B, N, heads, hd, scale, parts, p_drop, seed, site,
                          (cudaStream_t)stream);
  O2_FAIL(O2_ERR_ARG, "attn_bwd: unknown impl %d", impl);
}

int o2_attn_bwd_fused_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta,
                         float* dq_accum, int B, int N, int heads, int hd, float scale, int parts, float p_drop,
                         uint64_t seed, uint32_t site, cudaStream_t st);

extern "C" size_t o2_attn_bwd_fused_workspace(int B, int N, int heads, int hd) {
  return (size_t)B * heads * N * hd * sizeof(float);
}

extern "C" int o2_attn_bwd_fused(int parts, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                 float* delta, void* workspace, size_t ws_bytes, int B, int N, int heads, int hd, float scale,
                                 float p_drop, uint64_t seed, uint32_t site, void* stream) {
  O2_REQUIRE(B > 0 && N > 0 && heads > 0, "attn_bwd_fused: bad dims");
  O2_REQUIRE(workspace != nullptr && ws_bytes >= o2_attn_bwd_fused_workspace(B, N, heads, hd),
             "attn_bwd_fused: workspace of %zu bytes needed (o2_attn_bwd_fused_workspace), got %zu",
             o2_attn_bwd_fused_workspace(B, N, heads, hd), ws_bytes);
  return o2_attn_bwd_fused_tc(qkv, out, dout, lse, dqkv, delta, (float*)workspace, B, N, heads, hd, scale, parts, p_drop, seed,
                              site, (cudaStream_t)stream);
}

extern "C" int o2_attn_bwd_parts(int impl, int parts, const void* qkv, const void* out, const void* dout, const float* lse,
                                 void* dqkv, float* delta, int B, int N, int heads, int hd, float scale, void* stream) {
  return o2_attn_bwd_parts_drop(impl, parts, qkv, out, dout, lse, dqkv, delta, B, N, heads, hd, scale, 0.f, 0, 0, stream);
}

extern "C" int o2_attn_bwd(int impl, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                           float* delta, int B, int N, int heads, int hd, float scale, void* stream) {
  return o2_attn_bwd_parts(impl, O2_ATTN_BWD_ALL, qkv, out, dout, lse, dqkv, delta, B, N, heads, hd, scale, stream);
}
