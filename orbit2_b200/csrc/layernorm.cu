// LayerNorm forward / backward (HBM-bound).  reference: nn.LayerNorm(D, eps=1e-5) at vit_blocks.py:46,63 and
// res_slimvit.py:104; backward fused with the residual-stream gradient add of Block.forward (vit_blocks.py:78-79).
//
// One warp per token row (a group of G = 2 / 4 / 8 warps for the wide rows of interm_1b / 10b: D = 3072 / 8192), 16-byte
// vector loads, the row lives in registers between the statistics and the normalise pass (D <= kMaxVec*32*G*VEC);
// statistics reduced with warp shuffles (+ one shared-memory exchange when G > 1).  Backward = the same row kernel for dx plus a
// streaming column reduction for dgamma / dbeta (a fused register-accumulator version ran at 222 registers, one CTA
// per SM and 29 % of HBM peak).
#include "common.cuh"
#include <algorithm>
#include <type_traits>

namespace {

constexpr int kMaxVec = 8;   // vectors (16 B) per lane held in registers

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]); t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// sum over the G warps that share a row (G = 1: plain warp reduction).  Every thread of the CTA must call it.
template <int G>
__device__ __forceinline__ float group_sum(float v, float* sbuf) {
  v = warp_sum(v);
  if (G == 1) return v;
  const int warp = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sbuf[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < G; ++i) s += sbuf[(warp / G) * G + i];
  return s;
}

template <typename T, int NV, int G>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, T* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, long long rows,
                                                     int D, float eps) {
  constexpr int VN = Vec<T>::N;
  __shared__ float sbuf[8];
  const int lane = (threadIdx.x & (32 * G - 1));            // position inside the row's thread group
  long long row = (long long)blockIdx.x * (8 / G) + threadIdx.x / (32 * G);
  const bool valid = row < rows;
  if (G == 1 && !valid) return;
  if (!valid) row = rows - 1;                                // G > 1: keep every thread for the CTA barriers
  const int nvec = D / VN;
  const T* xr = x + row * D;
  float v[NV][VN];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32 * G;
    if (vi < nvec) {
      Vec<T>::load(xr + vi * VN, v[i]);
#pragma unroll
      for (int j = 0; j < VN; ++j) s += v[i][j];
    }
  }
  const float mu = group_sum<G>(s, sbuf) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32 * G;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < VN; ++j) { const float d = v[i][j] - mu; q += d * d; }
    }
  }
  const float rs = rsqrtf(group_sum<G>(q, sbuf) / D + eps);
  if (!valid) return;
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  T* yr = y + row * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32 * G;
    if (vi < nvec) {
      float o[VN], gm[VN], bt[VN];
#pragma unroll
      for (int j = 0; j < VN; j += 4) {
        Vec<float>::load(gamma + vi * VN + j, reinterpret_cast<float(&)[4]>(gm[j]));
        Vec<float>::load(beta + vi * VN + j, reinterpret_cast<float(&)[4]>(bt[j]));
      }
#pragma unroll
      for (int j = 0; j < VN; ++j) o[j] = (v[i][j] - mu) * rs * gm[j] + bt[j];
      Vec<T>::store(yr + vi * VN, o);
    }
  }
}

// dx = LN'(dy) (+ dres): one warp per row, same shape as the forward kernel (no column accumulators -> low register
// count, full occupancy).  The column gradients are a separate streaming reduction (ln_colgrad_kernel).
template <typename T, int NV, int G>
__global__ void __launch_bounds__(256) ln_bwd_dx_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const T* __restrict__ dres,
                                                        T* __restrict__ dx, long long rows, int D) {
  constexpr int VN = Vec<T>::N;
  __shared__ float sbuf[8];
  const int lane = (threadIdx.x & (32 * G - 1));
  long long row = (long long)blockIdx.x * (8 / G) + threadIdx.x / (32 * G);
  const bool valid = row < rows;
  if (G == 1 && !valid) return;
  if (!valid) row = rows - 1;
  const int nvec = D / VN;
  const float mu = mean[row], rs = rstd[row];
  float xh[NV][VN], gy[NV][VN];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32 * G;
    if (vi < nvec) {
      float xv[VN], dv[VN], gm[VN];
      Vec<T>::load(x + row * D + vi * VN, xv);
      Vec<T>::load(dy + row * D + vi * VN, dv);
#pragma unroll
      for (int j = 0; j < VN; j += 4) Vec<float>::load(gamma + vi * VN + j, reinterpret_cast<float(&)[4]>(gm[j]));
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        xh[i][j] = (xv[j] - mu) * rs;
        gy[i][j] = dv[j] * gm[j];
        s1 += gy[i][j];
        s2 += gy[i][j] * xh[i][j];
      }
    }
  }
  s1 = group_sum<G>(s1, sbuf) / D;
  s2 = group_sum<G>(s2, sbuf) / D;
  if (!valid) return;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32 * G;
    if (vi < nvec) {
      float o[VN];
      if (dres) Vec<T>::load(dres + row * D + vi * VN, o);
      else {
#pragma unroll
        for (int j = 0; j < VN; ++j) o[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < VN; ++j) o[j] += rs * (gy[i][j] - s1 - xh[i][j] * s2);
      Vec<T>::store(dx + row * D + vi * VN, o);
    }
  }
}

// Fused backward for rows that fit ONE 16-byte vector per lane (NV = 1: D = 1024 bf16 with G = 4 warps per row,
// D = 1024 fp32 with G = 8): dx and the column gradients in a single pass over dy / x (4 tensors moved instead of 6).
// A lane owns the same VN columns for every row it sees, so dgamma / dbeta partials are 2 VN registers; the CTA walks
// rows persistently, two per row-group per iteration (both rows' loads are in flight before the one barrier that
// exchanges the row statistics, double-buffered), and ends with a shared-memory reduction over its row groups and one
// atomic per column.
template <typename T, int G>
__global__ void __launch_bounds__(256) ln_bwd_fused_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                           const float* __restrict__ gamma, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const T* __restrict__ dres,
                                                           T* __restrict__ dx, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, long long rows, int D) {
  constexpr int VN = Vec<T>::N, RPC = 8 / G, R = 2;
  __shared__ float sred[2][8][R][2];
  __shared__ float scol[RPC > 1 ? RPC : 1][2][32 * G * VN];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & (32 * G - 1);
  const int grp = threadIdx.x / (32 * G);
  const bool col_ok = lane * VN < D;
  float gm[VN], ag[VN], ab[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) { gm[j] = 0.f; ag[j] = 0.f; ab[j] = 0.f; }
  if (col_ok) {
#pragma unroll
    for (int j = 0; j < VN; j += 4) Vec<float>::load(gamma + lane * VN + j, reinterpret_cast<float(&)[4]>(gm[j]));
  }
  const float invD = 1.0f / D;
  const long long stride = (long long)gridDim.x * RPC * R;
  int buf = 0;
  for (long long base = (long long)blockIdx.x * RPC * R; base < rows; base += stride) {
    float xh[R][VN], gy[R][VN], rs[R];
    long long row[R];
    bool ok[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      row[q] = base + q * RPC + grp;
      ok[q] = col_ok && row[q] < rows;
      float xv[VN], dv[VN];
#pragma unroll
      for (int j = 0; j < VN; ++j) { xv[j] = 0.f; dv[j] = 0.f; }
      float mu = 0.f;
      rs[q] = 0.f;
      if (ok[q]) {
        Vec<T>::load(x + row[q] * D + lane * VN, xv);
        Vec<T>::load(dy + row[q] * D + lane * VN, dv);
        mu = mean[row[q]];
        rs[q] = rstd[row[q]];
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        xh[q][j] = (xv[j] - mu) * rs[q];
        gy[q][j] = dv[j] * gm[j];
        s1 += gy[q][j];
        s2 = fmaf(gy[q][j], xh[q][j], s2);
        ag[j] = fmaf(dv[j], xh[q][j], ag[j]);
        ab[j] += dv[j];
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if ((threadIdx.x & 31) == 0) { sred[buf][warp][q][0] = s1; sred[buf][warp][q][1] = s2; }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < R; ++q) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < G; ++i) { s1 += sred[buf][grp * G + i][q][0]; s2 += sred[buf][grp * G + i][q][1]; }
      s1 *= invD;
      s2 *= invD;
      if (ok[q]) {
        float o[VN];
        if (dres) Vec<T>::load(dres + row[q] * D + lane * VN, o);
        else {
#pragma unroll
          for (int j = 0; j < VN; ++j) o[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < VN; ++j) o[j] += rs[q] * (gy[q][j] - s1 - xh[q][j] * s2);
        Vec<T>::store(dx + row[q] * D + lane * VN, o);
      }
    }
    buf ^= 1;
  }
  // column gradients: sum the CTA's row groups, one atomic per column
  if (RPC > 1) {
#pragma unroll
    for (int j = 0; j < VN; ++j) { scol[grp][0][lane * VN + j] = ag[j]; scol[grp][1][lane * VN + j] = ab[j]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * 32 * G * VN; i += 256) {
      const int which = i / (32 * G * VN), c = i % (32 * G * VN);
      float s_ = 0.f;
#pragma unroll
      for (int r = 0; r < RPC; ++r) s_ += scol[r][which][c];
      if (c < D) atomicAdd(which == 0 ? &dgamma[c] : &dbeta[c], s_);
    }
  } else if (col_ok) {
#pragma unroll
    for (int j = 0; j < VN; ++j) { atomicAdd(&dgamma[lane * VN + j], ag[j]); atomicAdd(&dbeta[lane * VN + j], ab[j]); }
  }
}

// Fused backward for bf16 rows of up to 1024 columns (the 117M / 8m widths): dx AND the column gradients in one pass
// over dy / x / dres (4 tensors moved instead of 6).  ONE WARP owns a row and keeps the dgamma / dbeta partials of its
// NV x 8 columns in fp32 registers for the whole persistent loop.  Rows arrive through a private 2-stage shared-memory
// ring filled by 1-D bulk copies (cp.async.bulk + mbarrier complete_tx): the bytes in flight live in shared memory, not
// in registers -- the register-resident versions of this kernel (12 warps per SM at 168 registers, or 4 warps per row
// with a CTA barrier) all stalled at 3.0 TB/s with ~48 KB in flight per SM; this one keeps 16 warps x 2 rows x 6 KB.
constexpr int kLnWarps = 8, kLnStages = 2;
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(ptx::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar)) : "memory");
}

// optional second output of the backward: the same dx through a token-stream dropout / drop-path mask (the gradient that
// enters the residual branch whose forward was x + drop_path(drop(branch)): mask function of dropout.cu, e = row * D + col)
struct LnDrop {
  __nv_bfloat16* out;           // nullptr: no second output
  uint32_t key, thr16;
  float inv_keep;
  const float* sample_scale;
  long long rows_per_sample;
  const uint64_t* step_word;
};

template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32, 2) ln_bwd_ring_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                       const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ dres,
                                                                       __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                                       float* __restrict__ dbeta, long long rows, int D, const LnDrop dm) {
  constexpr int W = NV * 256;                       // padded row width (elements)
  extern __shared__ __align__(128) uint8_t ln_smem[];
  __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(ln_smem);                 // [warp][stage][x | dy | dres][W]
  float* sg = reinterpret_cast<float*>(ring + kLnWarps * kLnStages * 3 * W);        // [W]
  float* scol = sg + W;                                                             // [2][W]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scol + 2 * W);                       // [warp][stage]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < W; i += kLnWarps * 32) { sg[i] = i < D ? gamma[i] : 0.f; scol[i] = 0.f; scol[W + i] = 0.f; }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kLnWarps * kLnStages; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int nvec = D >> 3;
  const uint32_t row_bytes = (uint32_t)D * 2u;
  const float invD = 1.0f / D;
  const long long nw = (long long)gridDim.x * kLnWarps;
  const long long row0 = (long long)blockIdx.x * kLnWarps + warp;
  __nv_bfloat16* my = ring + (size_t)warp * kLnStages * 3 * W;
  uint64_t* mybar = bars + warp * kLnStages;
  auto issue = [&](long long row, int st) {         // lane 0 only
    __nv_bfloat16* dst = my + (size_t)st * 3 * W;
    ptx::mbar_expect_tx(&mybar[st], (dres ? 3u : 2u) * row_bytes);
    bulk_g2s(dst, x + row * D, row_bytes, &mybar[st]);
    bulk_g2s(dst + W, dy + row * D, row_bytes, &mybar[st]);
    if (dres) bulk_g2s(dst + 2 * W, dres + row * D, row_bytes, &mybar[st]);
  };
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kLnStages; ++st)
      if (row0 + st * nw < rows) issue(row0 + st * nw, st);
  }
  float ag[NV][8], ab[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[i][j] = 0.f; ab[i][j] = 0.f; }
  int st = 0;
  uint32_t phase = 0;
  for (long long row = row0; row < rows; row += nw) {
    const float mu = mean[row], rs = rstd[row];
    ptx::mbar_wait(&mybar[st], phase);
    const __nv_bfloat16* sx = my + (size_t)st * 3 * W;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float xv[8], dv[8], gm[8];
        Vec<__nv_bfloat16>::load(sx + vi * 8, xv);
        Vec<__nv_bfloat16>::load(sx + W + vi * 8, dv);
        Vec<float>::load(sg + vi * 8, reinterpret_cast<float(&)[4]>(gm[0]));
        Vec<float>::load(sg + vi * 8 + 4, reinterpret_cast<float(&)[4]>(gm[4]));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mu) * rs, gy = dv[j] * gm[j];
          s1 += gy;
          s2 = fmaf(gy, xh, s2);
          ag[i][j] = fmaf(dv[j], xh, ag[i][j]);
          ab[i][j] += dv[j];
        }
      }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nvec) {
        float xv[8], dv[8], gm[8], o[8];
        Vec<__nv_bfloat16>::load(sx + vi * 8, xv);
        Vec<__nv_bfloat16>::load(sx + W + vi * 8, dv);
        Vec<float>::load(sg + vi * 8, reinterpret_cast<float(&)[4]>(gm[0]));
        Vec<float>::load(sg + vi * 8 + 4, reinterpret_cast<float(&)[4]>(gm[4]));
        if (dres) Vec<__nv_bfloat16>::load(sx + 2 * W + vi * 8, o);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mu) * rs, gy = dv[j] * gm[j];
          o[j] += rs * (gy - s1 - xh * s2);
        }
        Vec<__nv_bfloat16>::store(dx + row * D + vi * 8, o);
        if (dm.out) {
          float sc = dm.inv_keep;
          if (dm.sample_scale) sc *= __ldg(dm.sample_scale + row / dm.rows_per_sample);
          const uint32_t key = dm.key ^ ptx::step_word_mix(dm.step_word);
          const unsigned long long pair0 = (unsigned long long)(row * D + vi * 8) >> 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const unsigned long long pair = pair0 + j;
            const uint32_t h = ptx::lowbias32((uint32_t)pair ^ key ^ ((uint32_t)(pair >> 32) * 0x9E3779B1u));
            // the mask multiplies the ROUNDED dx, so the result is bit for bit what o2_dropout makes of the stored dx
            const float a = __bfloat162float(__float2bfloat16_rn(o[2 * j])), b = __bfloat162float(__float2bfloat16_rn(o[2 * j + 1]));
            o[2 * j] = ((h & 0xFFFFu) >= dm.thr16) ? a * sc : 0.f;
            o[2 * j + 1] = ((h >> 16) >= dm.thr16) ? b * sc : 0.f;
          }
          Vec<__nv_bfloat16>::store(dm.out + row * D + vi * 8, o);
        }
      }
    }
    __syncwarp();                                   // every lane is done reading this stage
    if (lane == 0 && row + kLnStages * nw < rows) {
      ptx::fence_proxy_async_smem();
      issue(row + kLnStages * nw, st);
    }
    if (++st == kLnStages) { st = 0; phase ^= 1; }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&scol[(lane + 32 * i) * 8 + j], ag[i][j]);
      atomicAdd(&scol[W + (lane + 32 * i) * 8 + j], ab[i][j]);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += kLnWarps * 32) { atomicAdd(&dgamma[c], scol[c]); atomicAdd(&dbeta[c], scol[W + c]); }
}

// dgamma[c] += sum_rows dy * (x - mean) * rstd, dbeta[c] += sum_rows dy: 32 column vectors x 8 row lanes per CTA,
// grid-strided over rows, smem transpose-reduce, one atomic per column per CTA.
template <typename T>
__global__ void __launch_bounds__(256) ln_colgrad_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                         long long rows, int D) {
  constexpr int VN = Vec<T>::N;
  __shared__ float sm[2][8][32 * VN + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * VN;
  float ag[VN], ab[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) { ag[j] = 0.f; ab[j] = 0.f; }
  if (col < D) {
    for (long long m = (long long)blockIdx.y * 8 + ty; m < rows; m += (long long)gridDim.y * 8) {
      const float mu = mean[m], rs = rstd[m];
      float xv[VN], dv[VN];
      Vec<T>::load(x + m * D + col, xv);
      Vec<T>::load(dy + m * D + col, dv);
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        ag[j] = fmaf(dv[j], (xv[j] - mu) * rs, ag[j]);
        ab[j] += dv[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VN; ++j) { sm[0][ty][tx * VN + j] = ag[j]; sm[1][ty][tx * VN + j] = ab[j]; }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 32 * VN; i += 256) {
    const int which = i / (32 * VN), c = i % (32 * VN);
    float s_ = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s_ += sm[which][r][c];
    const int gc = blockIdx.x * 32 * VN + c;
    if (gc < D) atomicAdd(which == 0 ? &dgamma[gc] : &dbeta[gc], s_);
  }
}

template <typename T>
int ln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, long long T_, int D,
           float eps, cudaStream_t st) {
  constexpr int VN = Vec<T>::N;
  const int nv = (D / VN + 31) / 32;                       // 16-byte vectors per lane if one warp held the row
#define O2_LN_FWD(NV, G)                                                                                       \
  ln_fwd_kernel<T, NV, G><<<(unsigned)((T_ + 8 / G - 1) / (8 / G)), 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, mean, \
                                                                                   rstd, T_, D, eps)
  if (nv <= 1) O2_LN_FWD(1, 1); else if (nv <= 2) O2_LN_FWD(2, 1); else if (nv <= 4) O2_LN_FWD(4, 1);
  else if (nv <= 8) O2_LN_FWD(8, 1); else if (nv <= 16) O2_LN_FWD(8, 2); else if (nv <= 32) O2_LN_FWD(8, 4);
  else O2_LN_FWD(8, 8);
#undef O2_LN_FWD
  O2_LAUNCH_CHECK();
  return O2_OK;
}

template <typename T>
int ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const void* dres,
           void* dx, float* dgamma, float* dbeta, long long T_, int D, cudaStream_t st, const LnDrop& dm) {
  constexpr int VN = Vec<T>::N;
  const int nv = (D / VN + 31) / 32;
  const bool ring = std::is_same<T, __nv_bfloat16>::value && nv <= 4 && !getenv("O2_LN_BWD_SPLIT");
  O2_REQUIRE(!dm.out || ring, "layernorm_bwd: the masked second output needs the bf16 row kernel (D <= 1024); apply o2_dropout to dx instead");
  if (ring) {
    auto launch = [&](auto kern, int nvt) -> int {
      const size_t smem = (size_t)kLnWarps * kLnStages * 3 * nvt * 256 * 2 + (size_t)3 * nvt * 256 * 4 + kLnWarps * kLnStages * 8;
      O2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLnWarps * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
      const long long grid = std::min((long long)o2_num_sms() * per_sm, (T_ + kLnWarps - 1) / kLnWarps);
      kern<<<(unsigned)grid, kLnWarps * 32, smem, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, gamma, mean, rstd,
                                                       (const __nv_bfloat16*)dres, (__nv_bfloat16*)dx, dgamma, dbeta, T_, D, dm);
      O2_LAUNCH_CHECK();
      return O2_OK;
    };
    if (nv <= 1) return launch(ln_bwd_ring_kernel<1>, 1);
    if (nv <= 2) return launch(ln_bwd_ring_kernel<2>, 2);
    return launch(ln_bwd_ring_kernel<4>, 4);
  }
  if (nv <= 8 && nv > 2 && !getenv("O2_LN_BWD_SPLIT")) {      // one vector per lane with G = 4 / 8 warps per row
    auto resident = [](const void* fn) {      // persistent grid = exactly the CTAs that are co-resident (one wave)
      int n = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, 256, 0) != cudaSuccess || n < 1) n = 1;
      return (long long)o2_num_sms() * n;
    };
    if (nv <= 4) {
      const long long want = resident((const void*)ln_bwd_fused_kernel<T, 4>);
      const long long per = 2 * 2, grid = std::min(want, (T_ + per - 1) / per);
      ln_bwd_fused_kernel<T, 4><<<(unsigned)grid, 256, 0, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres,
                                                              (T*)dx, dgamma, dbeta, T_, D);
    } else {
      const long long want = resident((const void*)ln_bwd_fused_kernel<T, 8>);
      const long long per = 1 * 2, grid = std::min(want, (T_ + per - 1) / per);
      ln_bwd_fused_kernel<T, 8><<<(unsigned)grid, 256, 0, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres,
                                                              (T*)dx, dgamma, dbeta, T_, D);
    }
    O2_LAUNCH_CHECK();
    return O2_OK;
  }
#define O2_LN_BWD(NV, G)                                                                                        \
  ln_bwd_dx_kernel<T, NV, G><<<(unsigned)((T_ + 8 / G - 1) / (8 / G)), 256, 0, st>>>(                            \
      (const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres, (T*)dx, T_, D)
  if (nv <= 1) O2_LN_BWD(1, 1); else if (nv <= 2) O2_LN_BWD(2, 1); else if (nv <= 4) O2_LN_BWD(4, 1);
  else if (nv <= 8) O2_LN_BWD(8, 1); else if (nv <= 16) O2_LN_BWD(8, 2); else if (nv <= 32) O2_LN_BWD(8, 4);
  else O2_LN_BWD(8, 8);
#undef O2_LN_BWD
  O2_LAUNCH_CHECK();
  const unsigned gx = (unsigned)((D + 32 * VN - 1) / (32 * VN));
  long long gy = ((long long)o2_num_sms() * 8) / gx;
  if (gy < 1) gy = 1;
  if (gy > (T_ + 7) / 8) gy = (T_ + 7) / 8;
  ln_colgrad_kernel<T><<<dim3(gx, (unsigned)gy), 256, 0, st>>>((const T*)dy, (const T*)x, mean, rstd, dgamma, dbeta, T_, D);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

int check_dims(long long T_, int D, int dtype) {
  O2_REQUIRE(T_ > 0 && D > 0, "layernorm: empty problem");
  O2_REQUIRE(dtype == O2_F32 || dtype == O2_BF16, "layernorm: bad dtype %d", dtype);
  const int vn = dtype == O2_F32 ? 4 : 8;
  O2_REQUIRE(D % vn == 0, "layernorm: D=%d must be a multiple of %d", D, vn);
  O2_REQUIRE(D / vn <= kMaxVec * 32 * 8, "layernorm: D=%d exceeds the register-resident limit %d", D, kMaxVec * 32 * 8 * vn);
  return O2_OK;
}

}  // namespace

extern "C" int o2_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                int64_t T_, int D, float eps, int dtype, void* stream) {
  int rc = check_dims(T_, D, dtype);
  if (rc) return rc;
  O2_REQUIRE(x && gamma && beta && y && mean && rstd, "layernorm_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == O2_F32 ? ln_fwd<float>(x, gamma, beta, y, mean, rstd, T_, D, eps, st)
                         : ln_fwd<__nv_bfloat16>(x, gamma, beta, y, mean, rstd, T_, D, eps, st);
}

extern "C" int o2_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                const void* dres, void* dx, float* dgamma, float* dbeta, int64_t T_, int D, int dtype,
                                void* stream) {
  int rc = check_dims(T_, D, dtype);
  if (rc) return rc;
  O2_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "layernorm_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  LnDrop dm;
  memset(&dm, 0, sizeof(dm));
  return dtype == O2_F32 ? ln_bwd<float>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, T_, D, st, dm)
                         : ln_bwd<__nv_bfloat16>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, T_, D, st, dm);
}

extern "C" int o2_layernorm_bwd_drop(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                     const void* dres, void* dx, void* dx_masked, float* dgamma, float* dbeta, int64_t T_, int D,
                                     const O2GemmDrop* drop, void* stream) {
  int rc = check_dims(T_, D, O2_BF16);
  if (rc) return rc;
  O2_REQUIRE(dy && x && gamma && mean && rstd && dx && dx_masked && dgamma && dbeta && drop, "layernorm_bwd_drop: null pointer");
  O2_REQUIRE(drop->p >= 0.f && drop->p < 1.f, "layernorm_bwd_drop: p=%f outside [0, 1)", (double)drop->p);
  O2_REQUIRE(!drop->sample_scale || drop->rows_per_sample > 0, "layernorm_bwd_drop: rows_per_sample must be > 0 with sample_scale");
  LnDrop dm;
  dm.out = (__nv_bfloat16*)dx_masked;
  dm.key = ptx::lowbias32((uint32_t)drop->seed ^ ptx::lowbias32(drop->site ^ (uint32_t)(drop->seed >> 32)));
  dm.thr16 = (uint32_t)floor((double)drop->p * 65536.0);
  dm.inv_keep = 1.f / (1.f - drop->p);
  dm.sample_scale = drop->sample_scale;
  dm.rows_per_sample = drop->sample_scale ? drop->rows_per_sample : 1;
  dm.step_word = o2_step_word();
  return ln_bwd<__nv_bfloat16>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, T_, D, (cudaStream_t)stream, dm);
}
