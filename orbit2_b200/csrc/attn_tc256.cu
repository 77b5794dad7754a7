// tcgen05 flash attention for head dim 256 (interm_10b: 32 heads x 256, configs/interm_10b.yaml:39-42), bf16 operands, fp32
// softmax / accumulation, sm_100a.  reference: components/attention.py:54-78 -- softmax(q k^T * hd^-0.5) v on the fused
// projection output qkv [B,N,3,heads,256]; out [B,N,heads,256]; lse [B,heads,N] (natural log); backward = autograd of that.
//
// One skeleton, four modes.  A CTA owns 128 rows ("owner" tile(s), resident in shared memory as the A operands) of one
// (batch, head) and streams the other sequence dimension in 64-row steps:
//     score GEMM(s)   S  [128 x 64] = A1 [128 x 256] T1^T        (SS-MMA, 16 x (128 x 64 x 16), fp32 in TMEM)
//                     dP [128 x 64] = A2 [128 x 256] T2^T        (DK / DQ only)
//     softmax threads X  [128 x 64] bf16, written over S in TMEM (thread = owner row = TMEM lane)
//     output GEMM     acc [128 x 256] += X Y                     (TS-MMA: A = X from TMEM, B = Y tile MN-major, N = 256)
//   mode | owner A1 / A2 | streamed T1, T2 |  X                                   |  Y  | acc
//   FWD  | Q             | K_j, V_j        |  P = 2^(S c - m)   (online, lazy max) |  T2 | O  -> out = O / l, lse
//   DV   | K             | Q_i, dO_i       |  P^T = 2^(S^T c - lse_i)              |  T2 | dV
//   DK   | K, V          | Q_i, dO_i       |  dS^T = P^T o (dP^T - delta_i)        |  T1 | dK (x scale)
//   DQ   | Q, dO         | K_j, V_j        |  dS = P o (dP - delta)                |  T1 | dQ (x scale)
// The backward is three launches (dV, dK, dQ: 8 GEMMs for 5 algorithmic ones): the dK / dV accumulators of one key tile
// alone fill the 512 TMEM columns (2 x 256), so a one-pass kernel cannot hold them next to S / dP.
// TMEM columns: S buffers [0,64) [64,128) | dP buffers [128,192) [192,256) | acc [256,512).  S / dP are double-buffered:
// the score GEMMs of step u + 1 run while the softmax threads work on step u.
// Shared memory: every operand tile is stored as four 64-column (128-byte rows, SWIZZLE_128B) atoms.  owner tiles
// 64 KiB each; T1 ring 2 x 32 KiB; T2 ring 2 x 32 KiB (FWD / DV: T2 is the late-consumed Y operand) or 1 x 32 KiB (DK / DQ:
// T2 only feeds the dP GEMM, which is issued first so that its stage is refilled under the S GEMM) -> 192 / 224 KiB.
// warps: 0 TMA producer | 1 tcgen05 issuer (warp-uniform loop, one elected lane) | 2-5 softmax + epilogue.
// Attention-probability dropout is not implemented here: p_drop > 0 at head dim 256 runs on the fp32 arm (ops.py).
#include "common.cuh"

namespace {

constexpr int kHD = 256;
constexpr int BO = 128;                       // owner rows per CTA
constexpr int BS = 64;                        // streamed rows per step
constexpr uint32_t kAtomO = BO * 128;         // 16 KiB: 128 rows x 64 columns of bf16
constexpr uint32_t kAtomS = BS * 128;         //  8 KiB
constexpr uint32_t kOwnerBytes = 4 * kAtomO;  // 64 KiB
constexpr uint32_t kStreamBytes = 4 * kAtomS; // 32 KiB
constexpr int kThreads = 192;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescale = 8.0f;

enum Mode { FWD = 0, DV = 1, DK = 2, DQ = 3 };

template <int MODE> struct Cfg {
  static constexpr int kOwners = (MODE == DK || MODE == DQ) ? 2 : 1;
  static constexpr int kT2Stages = (MODE == DK || MODE == DQ) ? 1 : 2;
  static constexpr bool kTwoScores = (MODE == DK || MODE == DQ);
  static constexpr uint32_t kSmem = kOwners * kOwnerBytes + (2 + kT2Stages) * kStreamBytes + 1024 /*stats*/ + 256 /*barriers*/ +
                                    1024 /*alignment slack*/;
};

struct Args256 {
  const __nv_bfloat16* qkv;   // [B, N, 3, heads, 256]
  __nv_bfloat16* out;         // FWD: [B, N, heads, 256]
  float* lse;                 // [B, heads, N]  (FWD writes, backward reads)
  const float* delta;         // [B, heads, N]
  __nv_bfloat16* dqkv;        // [B, N, 3, heads, 256]
  int B, N, heads;
  float scale, scale_log2;
};

__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
attn256_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do, const Args256 a) {
  using C = Cfg<MODE>;
  constexpr int kT2 = C::kT2Stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA1 = smem;
  uint8_t* sA2 = sA1 + kOwnerBytes;                           // DK / DQ only
  uint8_t* sT1 = sA1 + C::kOwners * kOwnerBytes;              // 2 stages
  uint8_t* sT2 = sT1 + 2 * kStreamBytes;                      // kT2 stages
  float* sStat = reinterpret_cast<float*>(sT2 + kT2 * kStreamBytes);   // [2 slots][2 (lse2, delta)][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + 1024);
  uint64_t* own_full = bars;            // 1
  uint64_t* t1_full = own_full + 1;     // 2
  uint64_t* t1_empty = t1_full + 2;     // 2
  uint64_t* t2_full = t1_empty + 2;     // 2
  uint64_t* t2_empty = t2_full + 2;     // 2
  uint64_t* s_full = t2_empty + 2;      // 2
  uint64_t* x_full = s_full + 2;        // 2 (128 arrivals)
  uint64_t* y_done = x_full + 2;        // 1: one phase per step
  uint64_t* acc_done = y_done + 1;      // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / a.heads, h = bh % a.heads;
  const int r0 = blockIdx.x * BO;                             // first owner row
  const int n_steps = (a.N + BS - 1) / BS;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::prefetch_tmap(&tmap_do);
    ptx::mbar_init(own_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&t1_full[s], 1);
      ptx::mbar_init(&t1_empty[s], 1);
      ptx::mbar_init(&t2_full[s], 1);
      ptx::mbar_init(&t2_empty[s], 1);
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&x_full[s], 128);
    }
    ptx::mbar_init(y_done, 1);
    ptx::mbar_init(acc_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColDP = 128, kColAcc = 256;

  // which slice (0 = q, 1 = k, 2 = v) of the fused projection each tile comes from; -1 = dO (second tensor map)
  constexpr int kA1 = (MODE == FWD || MODE == DQ) ? 0 : 1;
  constexpr int kA2 = (MODE == DK) ? 2 : -1;                  // DQ: dO
  constexpr int kS1 = (MODE == FWD || MODE == DQ) ? 1 : 0;    // K_j | Q_i
  constexpr int kS2 = (MODE == FWD || MODE == DQ) ? 2 : -1;   // V_j | dO_i

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      auto load64 = [&](uint8_t* dst, int which, int row, uint64_t* bar) {      // 64 rows x 256 columns -> 4 atoms of `stride`
        for (int c = 0; c < 4; ++c) {
          if (which >= 0) ptx::tma_load_4d(dst + c * kAtomS, &tmap_qkv, bar, c * 64, which * a.heads + h, row, b);
          else ptx::tma_load_4d(dst + c * kAtomS, &tmap_do, bar, c * 64, h, row, b);
        }
      };
      auto load_owner = [&](uint8_t* dst, int which) {                          // 128 rows: two 64-row boxes per atom
        for (int c = 0; c < 4; ++c)
          for (int hh = 0; hh < 2; ++hh) {
            if (which >= 0)
              ptx::tma_load_4d(dst + c * kAtomO + hh * kAtomS, &tmap_qkv, own_full, c * 64, which * a.heads + h, r0 + hh * 64, b);
            else
              ptx::tma_load_4d(dst + c * kAtomO + hh * kAtomS, &tmap_do, own_full, c * 64, h, r0 + hh * 64, b);
          }
      };
      ptx::mbar_expect_tx(own_full, C::kOwners * kOwnerBytes);
      load_owner(sA1, kA1);
      if (C::kOwners == 2) load_owner(sA2, kA2);
      for (int u = 0; u < n_steps; ++u) {
        const int s1 = u & 1, s2 = u % kT2;
        ptx::mbar_wait(&t1_empty[s1], ((u >> 1) & 1) ^ 1);
        ptx::mbar_expect_tx(&t1_full[s1], kStreamBytes);
        load64(sT1 + s1 * kStreamBytes, kS1, u * BS, &t1_full[s1]);
        ptx::mbar_wait(&t2_empty[s2], ((u / kT2) & 1) ^ 1);
        ptx::mbar_expect_tx(&t2_full[s2], kStreamBytes);
        load64(sT2 + s2 * kStreamBytes, kS2, u * BS, &t2_full[s2]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ tcgen05 issuer
    const uint32_t idesc_s = ptx::umma_idesc_bf16(BO, BS, 0, 0);        // scores: N = 64 streamed rows
    const uint32_t idesc_y = ptx::umma_idesc_bf16(BO, kHD, 0, 1);       // acc: A = X (TMEM, K-major), B = Y (MN-major), N = 256
    const uint64_t dA1 = ptx::umma_smem_desc(ptx::smem_u32(sA1), 16, 1024);
    const uint64_t dA2 = ptx::umma_smem_desc(ptx::smem_u32(sA2), 16, 1024);
    const uint64_t dT1 = ptx::umma_smem_desc(ptx::smem_u32(sT1), 16, 1024);
    const uint64_t dT2 = ptx::umma_smem_desc(ptx::smem_u32(sT2), 16, 1024);
    // Y: K = the 64 streamed rows in 8-row groups of 1024 bytes, N = 256 = four 64-column atoms kAtomS apart
    const uint64_t dY = ptx::umma_smem_desc(ptx::smem_u32(C::kTwoScores ? sT1 : sT2), kAtomS, 1024);
    auto issue_score = [&](int u) {
      const int buf = u & 1, s1 = u & 1, s2 = u % kT2;
      if (C::kTwoScores) {
        ptx::mbar_wait(&t2_full[s2], (u / kT2) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t d = tmem_base + kColDP + buf * BS;
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = desc_add(dA2, c * kAtomO + k * 32);
              const uint64_t db = desc_add(dT2, s2 * kStreamBytes + c * kAtomS + k * 32);
              if (c == 0 && k == 0) ptx::umma_ss_first(d, da, db, idesc_s);
              else ptx::umma_ss_acc(d, da, db, idesc_s);
            }
          ptx::umma_commit(&t2_empty[s2]);               // T2 only feeds this GEMM: refill it under the S GEMM
        }
        __syncwarp();
      }
      ptx::mbar_wait(&t1_full[s1], (u >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t d = tmem_base + buf * BS;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = desc_add(dA1, c * kAtomO + k * 32);
            const uint64_t db = desc_add(dT1, s1 * kStreamBytes + c * kAtomS + k * 32);
            if (c == 0 && k == 0) ptx::umma_ss_first(d, da, db, idesc_s);
            else ptx::umma_ss_acc(d, da, db, idesc_s);
          }
        if (!C::kTwoScores) ptx::umma_commit(&t1_empty[s1]);
        ptx::umma_commit(&s_full[buf]);
      }
      __syncwarp();
    };
    ptx::mbar_wait(own_full, 0);
    ptx::tc_fence_after();
    issue_score(0);
    for (int u = 0; u < n_steps; ++u) {
      const int buf = u & 1, s1 = u & 1, s2 = u % kT2;
      if (u + 1 < n_steps) issue_score(u + 1);
      if (!C::kTwoScores) ptx::mbar_wait(&t2_full[s2], (u / kT2) & 1);
      ptx::mbar_wait(&x_full[buf], (u >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t xa = tmem_base + buf * BS;           // packed bf16 X: 16 streamed rows = 8 columns per k-step
        const uint64_t yb = desc_add(dY, (C::kTwoScores ? s1 : s2) * kStreamBytes);
        if (u == 0) ptx::umma_ts(tmem_base + kColAcc, xa, yb, idesc_y, 0u);
        else ptx::umma_ts_acc(tmem_base + kColAcc, xa, yb, idesc_y);
#pragma unroll
        for (int k = 1; k < BS / 16; ++k) ptx::umma_ts_acc(tmem_base + kColAcc, xa + k * 8, desc_add(yb, k * 2048), idesc_y);
        if (C::kTwoScores) ptx::umma_commit(&t1_empty[s1]);
        else ptx::umma_commit(&t2_empty[s2]);
        ptx::umma_commit(y_done);
        if (u + 1 == n_steps) ptx::umma_commit(acc_done);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ softmax (thread = owner row = TMEM lane) + epilogue
    const int quarter = warp & 3;
    const int row_in = quarter * 32 + lane;
    const int row = r0 + row_in;
    const int tid_s = (warp - 2) * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float c = a.scale_log2;
    const size_t stat_base = ((size_t)b * a.heads + h) * a.N;
    float m_ref = -1e30f, l_sum = 0.f;                       // FWD
    float my_lse2 = 0.f, my_delta = 0.f;                     // DQ: per-row statistics
    if (MODE == DQ && row < a.N) {
      my_lse2 = a.lse[stat_base + row] * kLog2e;
      my_delta = a.delta[stat_base + row];
    }
    for (int u = 0; u < n_steps; ++u) {
      const int buf = u & 1;
      if (MODE == DV || MODE == DK) {                        // per-column statistics of the streamed queries -> slot u & 1
        if (tid_s < BS) {
          const int qi = u * BS + tid_s;
          float* st = sStat + buf * 128;
          st[tid_s] = qi < a.N ? a.lse[stat_base + qi] * kLog2e : 1e30f;          // 2^(s c - 1e30) = 0: rows past N
          st[64 + tid_s] = (MODE == DK && qi < a.N) ? a.delta[stat_base + qi] : 0.f;
        }
      }
      ptx::mbar_wait(&s_full[buf], (u >> 1) & 1);
      ptx::tc_fence_after();
      if (MODE == DV || MODE == DK) asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t s_addr = lane_addr + buf * BS;
      const uint32_t dp_addr = lane_addr + kColDP + buf * BS;
      if (MODE == FWD) {
        uint32_t s0[32], s1v[32];
        ptx::tmem_ld_32x32(s_addr, s0);
        ptx::tmem_ld_32x32(s_addr + 32, s1v);
        ptx::tmem_ld_wait();
        const int kbase = u * BS;
        const bool tail = kbase + BS > a.N;
        float mx = -1e30f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x0 = __uint_as_float(s0[i]) * c, x1 = __uint_as_float(s1v[i]) * c;
          if (tail) {
            if (kbase + i >= a.N) x0 = -1e30f;
            if (kbase + 32 + i >= a.N) x1 = -1e30f;
          }
          s0[i] = __float_as_uint(x0);
          s1v[i] = __float_as_uint(x1);
          mx = fmaxf(mx, fmaxf(x0, x1));
        }
        float alpha = 1.f;
        if (u == 0) {
          m_ref = mx;
        } else if (mx > m_ref + kRescale) {
          alpha = ptx::ex2(m_ref - mx);
          m_ref = mx;
        }
        if (__any_sync(0xffffffffu, alpha != 1.f)) {          // rare: O *= alpha for the rows whose maximum jumped
          ptx::mbar_wait(y_done, (u - 1) & 1);                // O += P V of step u - 1 has completed
          ptx::tc_fence_after();
          l_sum *= alpha;
#pragma unroll 1
          for (int cc = 0; cc < 8; ++cc) {
            uint32_t o[32];
            ptx::tmem_ld_32x32(lane_addr + kColAcc + cc * 32, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            ptx::tmem_st_32x32(lane_addr + kColAcc + cc * 32, o);
          }
          ptx::tmem_st_wait();
        }
        uint32_t pk[32];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ptx::ex2(__uint_as_float(s0[i]) - m_ref), p1 = ptx::ex2(__uint_as_float(s0[i + 1]) - m_ref);
          const float p2 = ptx::ex2(__uint_as_float(s1v[i]) - m_ref), p3 = ptx::ex2(__uint_as_float(s1v[i + 1]) - m_ref);
          sum += (p0 + p1) + (p2 + p3);
          pk[i >> 1] = pack_bf16x2(p0, p1);
          pk[16 + (i >> 1)] = pack_bf16x2(p2, p3);
        }
        l_sum += sum;
        ptx::tmem_st_32x32(s_addr, pk);
      } else {
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {                     // two 32-column chunks of the step
          uint32_t sv[32], dv[32];
          ptx::tmem_ld_32x32(s_addr + ch * 32, sv);
          if (MODE != DV) ptx::tmem_ld_32x32(dp_addr + ch * 32, dv);
          ptx::tmem_ld_wait();
          uint32_t pk[16];
          const float* st = sStat + buf * 128 + ch * 32;
          const int kbase = u * BS + ch * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0, p1;
            if (MODE == DQ) {
              p0 = ptx::ex2(fmaf(__uint_as_float(sv[i]), c, -my_lse2));
              p1 = ptx::ex2(fmaf(__uint_as_float(sv[i + 1]), c, -my_lse2));
              if (kbase + i >= a.N) p0 = 0.f;                // keys past N (zero-filled K rows give S = 0, not -inf)
              if (kbase + i + 1 >= a.N) p1 = 0.f;
              p0 *= __uint_as_float(dv[i]) - my_delta;
              p1 *= __uint_as_float(dv[i + 1]) - my_delta;
            } else {
              const float2 ls = *reinterpret_cast<const float2*>(st + i);
              p0 = ptx::ex2(fmaf(__uint_as_float(sv[i]), c, -ls.x));
              p1 = ptx::ex2(fmaf(__uint_as_float(sv[i + 1]), c, -ls.y));
              if (MODE == DK) {
                const float2 dl = *reinterpret_cast<const float2*>(st + 64 + i);
                p0 *= __uint_as_float(dv[i]) - dl.x;
                p1 *= __uint_as_float(dv[i + 1]) - dl.y;
              }
            }
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          // chunk 0 -> columns [0,16) of the S buffer, chunk 1 -> [16,32): chunk 1's S columns [32,64) are still intact
          ptx::tmem_st_32x16(s_addr + ch * 16, pk);
        }
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&x_full[buf]);
    }
    // ---- epilogue: acc -> bf16 rows of out / dqkv
    ptx::mbar_wait(acc_done, 0);
    ptx::tc_fence_after();
    float f;
    __nv_bfloat16* op;
    if (MODE == FWD) {
      f = 1.f / l_sum;
      op = a.out + (((size_t)b * a.N + row) * a.heads + h) * kHD;
      if (row < a.N) a.lse[stat_base + row] = kLn2 * (m_ref + log2f(l_sum));
    } else {
      constexpr int which = (MODE == DQ) ? 0 : (MODE == DK ? 1 : 2);
      f = (MODE == DV) ? 1.f : a.scale;
      op = a.dqkv + ((((size_t)b * a.N + row) * 3 + which) * a.heads + h) * kHD;
    }
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
      uint32_t o[32];
      ptx::tmem_ld_32x32(lane_addr + kColAcc + cc * 32, o);
      ptx::tmem_ld_wait();
      if (row < a.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[i]) * f, __uint_as_float(o[i + 1]) * f);
          w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * f, __uint_as_float(o[i + 3]) * f);
          w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * f, __uint_as_float(o[i + 5]) * f);
          w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * f, __uint_as_float(o[i + 7]) * f);
          *reinterpret_cast<uint4*>(op + cc * 32 + i) = w;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

// delta[b, h, n] = sum_d out[b, n, h, d] * dout[b, n, h, d]: one warp per 256-wide row (16 bytes per lane)
__global__ void attn256_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                     float* __restrict__ delta, int B, int N, int heads) {
  const long long gid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * N * heads;
  if (gid >= total) return;
  const uint4 o = *reinterpret_cast<const uint4*>(out + gid * kHD + lane * 8);
  const uint4 d = *reinterpret_cast<const uint4*>(dout + gid * kHD + lane * 8);
  const float2 o0 = unpack_bf16x2(o.x), o1 = unpack_bf16x2(o.y), o2 = unpack_bf16x2(o.z), o3 = unpack_bf16x2(o.w);
  const float2 d0 = unpack_bf16x2(d.x), d1 = unpack_bf16x2(d.y), d2 = unpack_bf16x2(d.z), d3 = unpack_bf16x2(d.w);
  float s = o0.x * d0.x + o0.y * d0.y + o1.x * d1.x + o1.y * d1.y + o2.x * d2.x + o2.y * d2.y + o3.x * d3.x + o3.y * d3.y;
  s = warp_sum(s);
  if (lane == 0) {
    const int h = (int)(gid % heads);
    const long long bn = gid / heads;
    delta[((size_t)(bn / N) * heads + h) * N + (bn % N)] = s;
  }
}

template <int MODE>
int launch256(const CUtensorMap& tq, const CUtensorMap& td, const Args256& a, cudaStream_t st) {
  O2_SET_SMEM_ONCE((attn256_kernel<MODE>), Cfg<MODE>::kSmem);
  dim3 grid((a.N + BO - 1) / BO, a.B * a.heads);
  attn256_kernel<MODE><<<grid, kThreads, Cfg<MODE>::kSmem, st>>>(tq, td, a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

int make_maps(CUtensorMap* tq, CUtensorMap* td, const void* qkv, const void* dout, int B, int N, int heads) {
  {
    uint64_t dims[4] = {(uint64_t)kHD, (uint64_t)(3 * heads), (uint64_t)N, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)kHD * 2, (uint64_t)3 * heads * kHD * 2, (uint64_t)N * 3 * heads * kHD * 2};
    uint32_t box[4] = {64, 1, (uint32_t)BS, 1};
    int rc = o2_make_tmap(tq, qkv, 2, 4, dims, str, box, 1);
    if (rc) return rc;
  }
  uint64_t dims[4] = {(uint64_t)kHD, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)kHD * 2, (uint64_t)heads * kHD * 2, (uint64_t)N * heads * kHD * 2};
  uint32_t box[4] = {64, 1, (uint32_t)BS, 1};
  return o2_make_tmap(td, dout ? dout : qkv, 2, 4, dims, str, box, 1);     // forward: no dO (the map is never used)
}

}  // namespace

int o2_attn_fwd_tc256(const void* qkv, void* out, float* lse, int B, int N, int heads, float scale, cudaStream_t st) {
  O2_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "attn_fwd_tc256: pointers must be 16-byte aligned");
  O2_REQUIRE((long long)B * heads <= 65535, "attn_fwd_tc256: B*heads too large");
  CUtensorMap tq, td;
  int rc = make_maps(&tq, &td, qkv, nullptr, B, N, heads);
  if (rc) return rc;
  Args256 a;
  memset(&a, 0, sizeof(a));
  a.qkv = (const __nv_bfloat16*)qkv; a.out = (__nv_bfloat16*)out; a.lse = lse;
  a.B = B; a.N = N; a.heads = heads; a.scale = scale; a.scale_log2 = scale * kLog2e;
  return launch256<FWD>(tq, td, a, st);
}

int o2_attn_bwd_tc256(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta, int B,
                      int N, int heads, float scale, int parts, cudaStream_t st) {
  O2_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)dqkv % 16) == 0,
             "attn_bwd_tc256: pointers must be 16-byte aligned");
  O2_REQUIRE((long long)B * heads <= 65535, "attn_bwd_tc256: B*heads too large");
  if (parts & O2_ATTN_BWD_DELTA) {
    const long long rows = (long long)B * N * heads;
    attn256_delta_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout,
                                                                             delta, B, N, heads);
    O2_LAUNCH_CHECK();
  }
  CUtensorMap tq, td;
  int rc = make_maps(&tq, &td, qkv, dout, B, N, heads);
  if (rc) return rc;
  Args256 a;
  memset(&a, 0, sizeof(a));
  a.qkv = (const __nv_bfloat16*)qkv; a.lse = const_cast<float*>(lse); a.delta = delta; a.dqkv = (__nv_bfloat16*)dqkv;
  a.B = B; a.N = N; a.heads = heads; a.scale = scale; a.scale_log2 = scale * kLog2e;
  if (parts & O2_ATTN_BWD_DKV) {
    rc = launch256<DV>(tq, td, a, st);
    if (rc) return rc;
    rc = launch256<DK>(tq, td, a, st);
    if (rc) return rc;
  }
  if (parts & O2_ATTN_BWD_DQ) return launch256<DQ>(tq, td, a, st);
  return O2_OK;
}
