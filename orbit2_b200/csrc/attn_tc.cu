// tcgen05 flash attention (bf16 operands, fp32 softmax/accumulation) for sm_100a.
// reference: components/attention.py:50-78 -- softmax(q k^T * hd^-0.5) v, bidirectional, no mask, on the fused
// projection output qkv [B,N,3,heads,hd]; out [B,N,heads,hd]; lse [B,heads,N] (natural log).
//
// Forward.  One CTA owns 256 query rows (two 128-row tiles) of one (batch, head) and streams the keys/values of that
// head in 128-row tiles; every K/V tile fetched by TMA is used by both query tiles.
//   warp 0      TMA producer: Q once, then a K ring and a V ring (cp.async.bulk.tensor, 128B swizzle)
//   warp 1      tcgen05 issuer (one elected lane): S_t = Q_t K_j^T into TMEM, O_t += P_t V_j with P_t read from TMEM
//   warps 2-5   softmax of query tile 0 (one thread per row = one TMEM lane)
//   warps 6-9   softmax of query tile 1
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384); P_t (bf16, 64 columns) overwrites the head
// of S_t in place, so the second GEMM never touches shared memory for its A operand.  While the softmax warps of
// one tile run exp2 on their S, the tensor core works on the other tile (ping-pong).  The running maximum is only
// refreshed -- and O rescaled, by the softmax thread that owns the row -- when it grew by more than 2^8, which keeps
// the exact result (P and the row sum share the same reference) and takes the O read-modify-write off the common path.
#include "common.cuh"

namespace {

constexpr int kHD = 64;
constexpr int BQ = 128;          // rows per query tile
constexpr int BKV = 128;         // keys per tile
constexpr int kThreadsF = 320;
constexpr int kThreadsB = 352;   // + one more issuing warp (one per tile)
constexpr uint32_t kTileBytes = BQ * kHD * 2;   // 16 KiB (Q, K and V tiles alike)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;
// One pair of exponentials in every kPoly* pairs is evaluated by the FMA-pipe polynomial instead of MUFU.EX2 (measured
// optima on B200 at N = 16200: forward 900 TFLOP/s @4 (with Q in TMEM; 857 @6, 808 @3), dq 1107 @4 vs 989, dkv best without: it is
// bound by its MMA issue chain, not by MUFU).
#ifndef O2_POLY_FWD
#define O2_POLY_FWD 4
#endif
#ifndef O2_POLY_DQ
#define O2_POLY_DQ 4
#endif
#ifndef O2_POLY_FUSED
#define O2_POLY_FUSED (1 << 20)
#endif
constexpr int kPolyFwd = O2_POLY_FWD, kPolyDq = O2_POLY_DQ, kPolyDkv = 1 << 20, kPolyFused = O2_POLY_FUSED;

#ifdef O2_TIMELINE
// Debug build only: CTA (0,0) records clock64() at the hand-off points of sub-tiles [kTlFirst, kTlFirst + kTlCount).
constexpr int kTlFirst = 100, kTlCount = 8, kTlSlots = 16;
__device__ long long g_timeline[kTlCount * 2 * kTlSlots];
#define O2_TL(u, t, slot)                                                                              \
  do {                                                                                                 \
    if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (u) >= kTlFirst && (u) < kTlFirst + kTlCount) \
      g_timeline[(((u)-kTlFirst) * 2 + (t)) * kTlSlots + (slot)] = clock64();                            \
  } while (0)
#else
#define O2_TL(u, t, slot) do { } while (0)
#endif

struct FwdArgs {
  const __nv_bfloat16* qkv;  // [B, N, 3, heads, hd] (the Q rows are read straight from here when Q lives in TMEM)
  __nv_bfloat16* out;
  float* lse;
  int B, N, heads, n_sub;   // n_sub = ceil(N / 64) key sub-tiles
  float scale_log2;         // softmax scale * log2(e)
  ptx::AttnDrop drop;       // kDrop kernels only
};

// attention-probability dropout, row-owner side (mask function: common.cuh attn_keep_word)
// row-owner kernels: AND-mask of packed pair PAIR (keys 2 PAIR, 2 PAIR + 1 of a 32-key block) from its keep word
template <int PAIR> __device__ __forceinline__ uint32_t drop_mask_bf16x2(uint32_t word) {
  return ptx::keep_mask_bf16x2<PAIR & 1>(word << (PAIR >> 1));
}
template <int PAIR = 0> __device__ __forceinline__ void drop_apply_bf16x2(uint32_t (&pk)[16], uint32_t word) {
  if constexpr (PAIR < 16) {
    pk[PAIR] &= drop_mask_bf16x2<PAIR>(word);
    drop_apply_bf16x2<PAIR + 1>(pk, word);
  }
}
// dP of the 32 keys of a block through the dropout mask (fp32 bit patterns): kept -> unchanged, dropped -> +0
template <int K = 0> __device__ __forceinline__ void drop_apply_f32(uint32_t (&d)[32], uint32_t word) {
  if constexpr (K < 32) {
    d[K] &= ptx::keep_mask_f32<(K >> 1) & 1, K & 1>(word << (K >> 2));
    drop_apply_f32<K + 1>(d, word);
  }
}

constexpr int BS = 64;            // key sub-tile (forward, dq) / query sub-tile (dkv) processed per MMA group
constexpr uint32_t kHalfBytes = BS * kHD * 2;   // 8 KiB: byte offset of rows 64.. inside a 128-row tile
template <int NH> struct FwdCfg {
  static constexpr int kStages = (NH == 1) ? 3 : 2;
  static constexpr uint32_t kSmemBytes = (2 + 2 * kStages) * NH * kTileBytes + 1024 + 256;
};

// descriptor arithmetic: the start-address field counts 16-byte units and smem addresses stay below 2^18, so a
// descriptor can be advanced by adding (bytes >> 4) to its low word.
__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// NH = head dim / 64: Q, K and V tiles are kept as NH separate 64-column (128-byte, one swizzle atom) half-tiles; S sums
// over them (NH x 4 MMAs), O has NH 64-column accumulators per query tile.  NH = 2 fills TMEM exactly (4 x 64 + 4 x 64).
template <int NH, bool kDrop>
__global__ void __launch_bounds__(kThreadsB, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  constexpr int kStagesF = FwdCfg<NH>::kStages;
  constexpr uint32_t kQKV = NH * kTileBytes;        // bytes of one Q / K / V tile (NH half-tiles)
  uint8_t* sQ = smem;                               // 2 tiles
  uint8_t* sK = sQ + 2 * kQKV;                      // kStagesF tiles
  uint8_t* sV = sK + kStagesF * kQKV;               // kStagesF tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStagesF * kQKV);
  uint64_t* q_full = bars;                          // 1
  uint64_t* k_full = q_full + 1;                    // kStagesF
  uint64_t* k_empty = k_full + kStagesF;
  uint64_t* v_full = k_empty + kStagesF;
  uint64_t* v_empty = v_full + kStagesF;
  uint64_t* s_full = v_empty + kStagesF;            // [tile][buffer] = 4
  uint64_t* p_full = s_full + 4;                    // [tile][buffer] = 4 (128 arrivals)
  uint64_t* o_done = p_full + 4;                    // 2: one phase per sub-tile (gates the lazy rescale)
  uint64_t* o_final = o_done + 2;                   // 2: completes once, after the last P V of the tile
  uint64_t* q_ready = o_final + 2;                  // 2: Q_t has been written to TMEM (128 arrivals; kQTmem only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 2);
  // Head dim 64: Q_t (bf16, 32 packed columns) sits in the 128 TMEM columns that S / O leave free and is the A operand of
  // S = Q K^T from there (TS-MMA), exactly like P for P V.  An SS-MMA of 128 x 64 x 16 reads 4 KiB of A + 2 KiB of B per
  // instruction and runs at the shared-memory bandwidth (58 cycles instead of 32); with the constant operand in TMEM
  // only the 2 KiB of K are read.  Head dim 128 has no spare columns and keeps Q in shared memory.
  constexpr bool kQTmem = (NH == 1);
  constexpr uint32_t kQCol = 384;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / a.heads, h = bh % a.heads;
  const int q0 = blockIdx.x * 2 * BQ;
  const int n_sub = a.n_sub;
  const int n_kv = (n_sub + 1) >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < kStagesF; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], 2);             // one commit per issuing warp
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], 2);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_full[i], 128);
    }
    for (int t = 0; t < 2; ++t) {
      ptx::mbar_init(&o_done[t], 1);
      ptx::mbar_init(&o_final[t], 1);
      ptx::mbar_init(&q_ready[t], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S buffer (t, bb) at (2t + bb) * 64 (64 fp32 columns; bf16 P overwrites its first 32), O_t at 256 + 64t

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (!kQTmem) {
        ptx::mbar_expect_tx(q_full, 2 * kQKV);
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          ptx::tma_load_4d(sQ + hh * kTileBytes, &tmap_qkv, q_full, hh * 64, h, q0, b);
          ptx::tma_load_4d(sQ + kQKV + hh * kTileBytes, &tmap_qkv, q_full, hh * 64, h, q0 + BQ, b);
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        ptx::mbar_wait(&k_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&k_full[stage], kQKV);
#pragma unroll
        for (int hh = 0; hh < NH; ++hh)
          ptx::tma_load_4d(sK + stage * kQKV + hh * kTileBytes, &tmap_qkv, &k_full[stage], hh * 64, a.heads + h, j * BKV, b);
        ptx::mbar_wait(&v_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&v_full[stage], kQKV);
#pragma unroll
        for (int hh = 0; hh < NH; ++hh)
          ptx::tma_load_4d(sV + stage * kQKV + hh * kTileBytes, &tmap_qkv, &v_full[stage], hh * 64, 2 * a.heads + h, j * BKV, b);
        if (++stage == kStagesF) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ------------------------------------------------------------ tcgen05 issuers, one warp per query tile
    // The whole warp runs this loop (warp-uniform control flow, so descriptors live in uniform registers and an MMA
    // costs a couple of issue slots); one elected lane executes the tcgen05 instructions.  Separate issuers keep the
    // two tiles' softmax -> MMA -> softmax chains from blocking each other; the tensor pipe interleaves them.
    {
      const int t = (warp == 1) ? 0 : 1;
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BS, 0, 0);
      const uint32_t idesc_o = ptx::umma_idesc_bf16(BQ, kHD, 0, 1);   // A = P (TMEM, K-major), B = V (MN-major)
      const uint64_t dq0 = ptx::umma_smem_desc(ptx::smem_u32(sQ), 16, 1024);
      const uint64_t dk0 = ptx::umma_smem_desc(ptx::smem_u32(sK), 16, 1024);
      const uint64_t dv0 = ptx::umma_smem_desc(ptx::smem_u32(sV), 8192, 1024);
      // S_t(u) = Q_t K_u^T into buffer (t, u & 1); K sub-tile u = rows [64 (u&1), +64) of K tile u >> 1
      auto issue_s = [&](int t, int u, int kstage) {
        if (ptx::elect_one()) {
          const uint64_t kd = desc_add(dk0, kstage * kQKV + (u & 1) * kHalfBytes);
          const uint32_t d = tmem_base + (2 * t + (u & 1)) * BS;
          if (kQTmem) {
            const uint32_t qa = tmem_base + kQCol + t * 32;      // 8 packed columns per 16-wide K slice
            ptx::umma_ts(d, qa, kd, idesc_s, 0u);
#pragma unroll
            for (int k = 1; k < 4; ++k) ptx::umma_ts_acc(d, qa + k * 8, desc_add(kd, k * 32), idesc_s);
          } else {
            const uint64_t qd = desc_add(dq0, t * kQKV);
            ptx::umma_ss_first(d, qd, kd, idesc_s);
#pragma unroll
            for (int k = 1; k < NH * 4; ++k)          // 16-wide K slices: 4 per 64-column half-tile
              ptx::umma_ss_acc(d, desc_add(qd, (k >> 2) * kTileBytes + (k & 3) * 32), desc_add(kd, (k >> 2) * kTileBytes + (k & 3) * 32),
                               idesc_s);
          }
          ptx::umma_commit(&s_full[2 * t + (u & 1)]);
        }
        __syncwarp();
      };
      if (kQTmem) ptx::mbar_wait(&q_ready[t], 0);
      else ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&k_full[0], 0);
      ptx::tc_fence_after();
      issue_s(t, 0, 0);
      if (n_sub > 1) issue_s(t, 1, 0);
      if (ptx::elect_one()) ptx::umma_commit(&k_empty[0]);
      __syncwarp();
      int kstage = 1 % kStagesF;               // stage / phase of K tile (u + 2) >> 1 while u runs over tile j
      uint32_t kphase = (kStagesF == 1) ? 1 : 0;
      int vstage = 0;
      uint32_t vphase = 0;
      for (int u = 0; u < n_sub; ++u) {
        const int half = u & 1;
        const bool more = (u + 2 < n_sub);
        if (half == 0) {
          ptx::mbar_wait(&v_full[vstage], vphase);
          ptx::tc_fence_after();
        }
        const uint64_t vd = desc_add(dv0, vstage * kQKV + half * kHalfBytes);
        {
          O2_TL(u, t, 0);
          ptx::mbar_wait(&p_full[2 * t + half], (u >> 1) & 1);
          ptx::tc_fence_after();
          O2_TL(u, t, 1);
          if (ptx::elect_one()) {
            const uint32_t pa = tmem_base + (2 * t + half) * BS;
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {          // O_t[:, 64 hh .. +64) += P V[:, 64 hh .. +64)
              const uint32_t od = tmem_base + 256 + (t * NH + hh) * 64;
              const uint64_t vh = desc_add(vd, hh * kTileBytes);
              ptx::umma_ts(od, pa, vh, idesc_o, u > 0 ? 1u : 0u);
#pragma unroll
              for (int k = 1; k < BS / 16; ++k) ptx::umma_ts_acc(od, pa + k * 8, desc_add(vh, k * 2048), idesc_o);
            }
            ptx::umma_commit(&o_done[t]);
            if (u + 1 == n_sub) ptx::umma_commit(&o_final[t]);
          }
          __syncwarp();
          O2_TL(u, t, 3);
          if (more) {
            if (half == 0) {
              ptx::mbar_wait(&k_full[kstage], kphase);
              ptx::tc_fence_after();
            }
            O2_TL(u, t, 9);
            issue_s(t, u + 2, kstage);
            O2_TL(u, t, 2);
            if (half == 1 || u + 3 >= n_sub) {            // this warp's last S that reads the K tile has been issued
              if (ptx::elect_one()) ptx::umma_commit(&k_empty[kstage]);
              __syncwarp();
              if (++kstage == kStagesF) { kstage = 0; kphase ^= 1; }
            }
          }
        }
        if (half == 1 || u + 1 == n_sub) {
          if (ptx::elect_one()) ptx::umma_commit(&v_empty[vstage]);
          __syncwarp();
          if (++vstage == kStagesF) { vstage = 0; vphase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / correction / epilogue
    const int t = (warp - 2) >> 2;          // query tile of this warp group
    const int quarter = warp & 3;           // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;      // row inside the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t o_addr = lane_addr + 256 + t * NH * 64;
    const float sc = a.scale_log2;
    // attention dropout: P V uses the masked probabilities, the normaliser l the unmasked ones (attention.py:73-76)
    const uint32_t drop_key = kDrop ? ptx::attn_drop_key(a.drop, bh) : 0u;
    const uint32_t drop_q = (uint32_t)(q0 + t * BQ + r);
    if (kQTmem) {                           // this thread's query row -> its TMEM lane, 64 bf16 = 32 packed columns
      const int qrow = q0 + t * BQ + r;
      uint32_t qw[32];
      if (qrow < a.N) {
        const uint4* src = reinterpret_cast<const uint4*>(a.qkv + (((size_t)b * a.N + qrow) * 3 * a.heads + h) * kHD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 v = src[i];
          qw[4 * i] = v.x; qw[4 * i + 1] = v.y; qw[4 * i + 2] = v.z; qw[4 * i + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) qw[i] = 0u;
      }
      ptx::tmem_st_32x32(lane_addr + kQCol + t * 32, qw);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&q_ready[t]);
    }
    float m_ref = -INFINITY;
    uint64_t lA = 0ull, lB = 0ull;          // packed partial row sums
    const int tail = a.N - (n_sub - 1) * BS;    // valid keys in the last sub-tile (1..64)
    for (int u = 0; u < n_sub; ++u) {
      const int bb = u & 1;
      const uint32_t s_addr = lane_addr + (2 * t + bb) * BS;
      if (quarter == 2) O2_TL(u, t, 4);
      ptx::mbar_wait(&s_full[2 * t + bb], (u >> 1) & 1);
      ptx::tc_fence_after();
      if (quarter == 2) O2_TL(u, t, 5);
      uint32_t v0[32], v1[32];
      ptx::tmem_ld_32x32(s_addr, v0);
      ptx::tmem_ld_32x32(s_addr + 32, v1);
      ptx::tmem_ld_wait();
      if (quarter == 2) O2_TL(u, t, 6);
      const bool ragged = (u == n_sub - 1) && (tail != BS);
      if (ragged) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= tail) v0[i] = 0xff800000u;          // -inf
          if (32 + i >= tail) v1[i] = 0xff800000u;
        }
      }
      // 3-input max (FMNMX3): half the issue slots of a 2-input chain
      float mxa = ptx::max3(__uint_as_float(v0[0]), __uint_as_float(v0[1]), __uint_as_float(v0[2]));
      float mxb = ptx::max3(__uint_as_float(v1[0]), __uint_as_float(v1[1]), __uint_as_float(v1[2]));
#pragma unroll
      for (int i = 3; i < 31; i += 2) {
        mxa = ptx::max3(mxa, __uint_as_float(v0[i]), __uint_as_float(v0[i + 1]));
        mxb = ptx::max3(mxb, __uint_as_float(v1[i]), __uint_as_float(v1[i + 1]));
      }
      mxa = fmaxf(mxa, __uint_as_float(v0[31]));
      mxb = fmaxf(mxb, __uint_as_float(v1[31]));
      const float m_new = fmaxf(m_ref, fmaxf(mxa, mxb) * sc);
      const bool need = __any_sync(0xffffffffu, m_new - m_ref > kRescaleThreshold);
      if (need) {
        const float f = ptx::ex2(m_ref - m_new);     // 0 on the first sub-tile (m_ref = -inf)
        lA = ptx::mul2(lA, ptx::pack2(f, f));
        lB = ptx::mul2(lB, ptx::pack2(f, f));
        m_ref = m_new;
        if (u > 0) {
          ptx::mbar_wait(&o_done[t], (u - 1) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < NH * 2; ++c) {
            uint32_t w[32];
            ptx::tmem_ld_32x32(o_addr + c * 32, w);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) * f);
            ptx::tmem_st_32x32(o_addr + c * 32, w);
          }
        }
      }
      // exp2 in packed pairs (FFMA2 / FADD2); every kPolyFwd-th pair is evaluated by the FMA-pipe polynomial
      // instead of MUFU.EX2 (the MUFU pipe is the softmax bottleneck at head dim 64)
      const uint64_t sc2 = ptx::pack2(sc, sc), nm2 = ptx::pack2(-m_ref, -m_ref);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const uint64_t X = ptx::fma2(ptx::pack2u(v0[i], v0[i + 1]), sc2, nm2);
        const uint64_t Pq = (((i >> 1) % kPolyFwd) == kPolyFwd - 1) ? ptx::exp2_pair<true>(X) : ptx::exp2_pair<false>(X);
        if ((i >> 1) & 1) lB = ptx::add2(lB, Pq); else lA = ptx::add2(lA, Pq);
        pk[i >> 1] = ptx::pack_bf16x2_pair(Pq);
      }
      if (kDrop) drop_apply_bf16x2(pk, ptx::attn_keep_word(a.drop, drop_key, drop_q, (uint32_t)(2 * u)));
      ptx::tmem_st_32x16(s_addr, pk);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const uint64_t X = ptx::fma2(ptx::pack2u(v1[i], v1[i + 1]), sc2, nm2);
        const uint64_t Pq = (((i >> 1) % kPolyFwd) == kPolyFwd - 1) ? ptx::exp2_pair<true>(X) : ptx::exp2_pair<false>(X);
        if ((i >> 1) & 1) lB = ptx::add2(lB, Pq); else lA = ptx::add2(lA, Pq);
        pk[i >> 1] = ptx::pack_bf16x2_pair(Pq);
      }
      if (kDrop) drop_apply_bf16x2(pk, ptx::attn_keep_word(a.drop, drop_key, drop_q, (uint32_t)(2 * u + 1)));
      ptx::tmem_st_32x16(s_addr + 16, pk);
      if (quarter == 2) O2_TL(u, t, 7);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[2 * t + bb]);
      if (quarter == 2) O2_TL(u, t, 8);
    }
    // epilogue: O / l -> bf16, lse
    // (o_done may be up to two phases behind here, which a parity wait cannot tell apart: separate barrier)
    ptx::mbar_wait(&o_final[t], 0);
    ptx::tc_fence_after();
    const int row = q0 + t * BQ + r;
    float l0, l1, l2, l3;
    ptx::unpack2(lA, l0, l1);
    ptx::unpack2(lB, l2, l3);
    const float l = (l0 + l1) + (l2 + l3);
    const float inv = (kDrop ? a.drop.inv_keep : 1.f) / l;
    __nv_bfloat16* op = a.out + (((size_t)b * a.N + row) * a.heads + h) * (NH * 64);
#pragma unroll 1
    for (int c = 0; c < NH * 2; ++c) {
      uint32_t o[32];
      ptx::tmem_ld_32x32(o_addr + c * 32, o);
      ptx::tmem_ld_wait();
      if (row < a.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c * 32 + i) = w;
        }
      }
    }
    if (row < a.N) {
      a.lse[((size_t)b * a.heads + h) * a.N + row] = (m_ref + log2f(l)) * kLn2;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

// =====================================================================================================
// Backward.  Two kernels so that every gradient is accumulated in TMEM by exactly one CTA (no atomics,
// deterministic): with hd = 64 a single-pass design would have to push 32 KiB of fp32 dQ partials per
// 128x128 tile pair through L2 atomics (~7 TB/s at tensor-core speed), which is slower than recomputing S.
//   preprocess : delta[b,h,n] = sum_d dO * O (fp32), lse and delta stored pre-scaled for exp2
//   dq kernel  : CTA owns 256 query rows; per 64-key sub-tile  S = Q K^T, dP = dO V^T (TMEM) ->
//                dS = P o (dP - delta) (bf16, over S in TMEM) -> dQ += dS K  (TS-MMA, K as MN-major B)
//   dkv kernel : CTA owns 256 keys; per 64-query sub-tile  S^T = K Q^T, dP^T = V dO^T (TMEM) ->
//                P^T over S^T, dS^T over dP^T (bf16) -> dV += P^T dO, dK += dS^T Q  (TS-MMA)
// The softmax scale is applied once to dQ / dK in the epilogue.
// =====================================================================================================
// NH = head dim / 64 (as in the forward kernel).  Head dim 64: two 128-row tiles per CTA in ping-pong, 3-stage rings.
// Head dim 128 (interm_1b): dK / dV (dQ) need 2 x 128 TMEM columns per tile and every operand tile is 32 KiB, so a CTA
// owns ONE tile and the rings have 2 stages.
template <int NH> struct BwdCfg {
  static constexpr int kTiles = (NH == 1) ? 2 : 1;
  static constexpr int kStages = (NH == 1) ? 3 : 2;
  static constexpr uint32_t kT = NH * kTileBytes;                     // bytes of one 128-row operand tile
  static constexpr uint32_t kDqSmem = (2 * kTiles + 2 * kStages) * kT + 1024 + 4 * 1024 + 256;
  static constexpr uint32_t kDkvSmem = (2 * kTiles + 2 * kStages) * kT + 2 * kTileBytes + 1024 + 256;
  static constexpr int kDqThreads = (kTiles == 2) ? kThreadsB : 192;  // producer + issuer + 4 softmax warps per tile (+ issuer)
  static constexpr int kTmemTile = 256;                               // TMEM column stride between the two tiles (NH = 1)
};

struct BwdArgs {
  const __nv_bfloat16* qkv;   // [B, N, 3, heads, hd] (K / V rows are copied straight from here into TMEM by the v2 dK/dV kernel)
  const float* lse;       // [B, heads, N] natural log
  const float* delta;     // [B, heads, N]
  __nv_bfloat16* dqkv;    // [B, N, 3, heads, hd]
  int B, N, heads, n_sub; // n_sub = ceil(N / 64)
  float scale, scale_log2;
  ptx::AttnDrop drop;     // kDrop kernels only
};

template <int NH>
__global__ void attn_delta_bf16_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                       float* __restrict__ delta, int B, int N, int heads) {
  // one (8 NH)-thread group per (b, n, h) row of 64 NH bf16 (16 B per thread)
  constexpr int G = 8 * NH;
  const long long gid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
  const int sub = threadIdx.x & (G - 1);
  const long long total = (long long)B * N * heads;
  float s = 0.f;
  if (gid < total) {
    const uint4 o = *reinterpret_cast<const uint4*>(out + gid * (NH * 64) + sub * 8);
    const uint4 d = *reinterpret_cast<const uint4*>(dout + gid * (NH * 64) + sub * 8);
    const float2 o0 = unpack_bf16x2(o.x), o1 = unpack_bf16x2(o.y), o2 = unpack_bf16x2(o.z), o3 = unpack_bf16x2(o.w);
    const float2 d0 = unpack_bf16x2(d.x), d1 = unpack_bf16x2(d.y), d2 = unpack_bf16x2(d.z), d3 = unpack_bf16x2(d.w);
    s = o0.x * d0.x + o0.y * d0.y + o1.x * d1.x + o1.y * d1.y + o2.x * d2.x + o2.y * d2.y + o3.x * d3.x + o3.y * d3.y;
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (NH == 2) s += __shfl_xor_sync(0xffffffffu, s, 8);
  if (gid < total && sub == 0) {
    const int h = (int)(gid % heads);
    const long long bn = gid / heads;
    const int n = (int)(bn % N);
    const int b = (int)(bn / N);
    delta[((size_t)b * heads + h) * N + n] = s;
  }
}


// ---------------------------------------------------------------------------------------------- dQ
// TMEM columns of query tile t (base t*256): S buffers [0,64) / [64,128), dP [128,192), dQ [192, 192 + 64 NH).
// S is double-buffered (S(u+2) is issued as soon as dQ(u) has consumed the dS written over S(u)); dP has one buffer
// that the softmax warps release as soon as they have it in registers, so dP(u+1) is computed while they work on u.
template <int NH, bool kDrop>
__global__ void __launch_bounds__(BwdCfg<NH>::kDqThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                   const BwdArgs a) {
  constexpr int kStagesB = BwdCfg<NH>::kStages;
  constexpr int kTiles = BwdCfg<NH>::kTiles;
  constexpr uint32_t kT = BwdCfg<NH>::kT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* sQ = smem;                               // kTiles tiles
  uint8_t* sdO = sQ + kTiles * kT;                  // kTiles tiles
  uint8_t* sK = sdO + kTiles * kT;                  // kStagesB tiles
  uint8_t* sV = sK + kStagesB * kT;                 // kStagesB tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStagesB * kT + 4 * 1024);
  uint64_t* q_full = bars;                          // 1
  uint64_t* kv_full = q_full + 1;                   // kStagesB
  uint64_t* kv_empty = kv_full + kStagesB;
  uint64_t* s_full = kv_empty + kStagesB;           // [tile][buffer] = 4
  uint64_t* ds_full = s_full + 4;                   // [tile][buffer] = 4 (128 arrivals): dS written over S
  uint64_t* dp_full = ds_full + 4;                  // 2
  uint64_t* dp_free = dp_full + 2;                  // 2 (128 arrivals): dP is in registers
  uint64_t* dq_final = dp_free + 2;                 // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_final + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / a.heads, h = bh % a.heads;
  const int q0 = blockIdx.x * kTiles * BQ;
  const int n_sub = a.n_sub;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::prefetch_tmap(&tmap_do);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < kStagesB; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], kTiles);       // one commit per issuing warp
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&ds_full[i], 128);
    }
    for (int t = 0; t < 2; ++t) {
      ptx::mbar_init(&dp_full[t], 1);
      ptx::mbar_init(&dp_free[t], 128);
      ptx::mbar_init(&dq_final[t], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, 2 * kTiles * kT);
#pragma unroll
      for (int t = 0; t < kTiles; ++t)
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          ptx::tma_load_4d(sQ + t * kT + hh * kTileBytes, &tmap_qkv, q_full, hh * 64, h, q0 + t * BQ, b);
          ptx::tma_load_4d(sdO + t * kT + hh * kTileBytes, &tmap_do, q_full, hh * 64, h, q0 + t * BQ, b);
        }
      int stage = 0;
      uint32_t phase = 0;
      const int n_kv = (n_sub + 1) / 2;
      for (int j = 0; j < n_kv; ++j) {
        ptx::mbar_wait(&kv_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&kv_full[stage], 2 * kT);
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          ptx::tma_load_4d(sK + stage * kT + hh * kTileBytes, &tmap_qkv, &kv_full[stage], hh * 64, a.heads + h, j * BKV, b);
          ptx::tma_load_4d(sV + stage * kT + hh * kTileBytes, &tmap_qkv, &kv_full[stage], hh * 64, 2 * a.heads + h, j * BKV, b);
        }
        if (++stage == kStagesB) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 || (kTiles == 2 && warp == 10)) {
    {   // one issuing warp per query tile (warp-uniform loop, see the forward kernel): the two tiles' dependency
        // chains never block each other; the tensor pipe interleaves their MMAs
      const int t = (warp == 1) ? 0 : 1;
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BS, 0, 0);    // S / dP: N = 64 keys
      const uint32_t idesc_q = ptx::umma_idesc_bf16(BQ, kHD, 0, 1);   // dQ: A = dS (TMEM), B = K (MN-major)
      const uint64_t dq0 = ptx::umma_smem_desc(ptx::smem_u32(sQ), 16, 1024);
      const uint64_t ddo0 = ptx::umma_smem_desc(ptx::smem_u32(sdO), 16, 1024);
      const uint64_t dk0 = ptx::umma_smem_desc(ptx::smem_u32(sK), 16, 1024);
      const uint64_t dv0 = ptx::umma_smem_desc(ptx::smem_u32(sV), 16, 1024);
      const uint64_t dkm0 = ptx::umma_smem_desc(ptx::smem_u32(sK), 8192, 1024);   // K as MN-major B
      auto issue_s = [&](int t, int u, int stage) {
        if (ptx::elect_one()) {
          const uint64_t qa = desc_add(dq0, t * kT);
          const uint64_t ka = desc_add(dk0, stage * kT + (u & 1) * kHalfBytes);
          const uint32_t d = tmem_base + t * 256 + (u & 1) * BS;
          ptx::umma_ss_first(d, qa, ka, idesc_s);
#pragma unroll
          for (int k = 1; k < NH * 4; ++k)
            ptx::umma_ss_acc(d, desc_add(qa, (k >> 2) * kTileBytes + (k & 3) * 32),
                             desc_add(ka, (k >> 2) * kTileBytes + (k & 3) * 32), idesc_s);
          ptx::umma_commit(&s_full[2 * t + (u & 1)]);
        }
        __syncwarp();
      };
      auto issue_dp = [&](int t, int u, int stage) {
        if (ptx::elect_one()) {
          const uint64_t da = desc_add(ddo0, t * kT);
          const uint64_t va = desc_add(dv0, stage * kT + (u & 1) * kHalfBytes);
          const uint32_t d = tmem_base + t * 256 + 128;
          ptx::umma_ss_first(d, da, va, idesc_s);
#pragma unroll
          for (int k = 1; k < NH * 4; ++k)
            ptx::umma_ss_acc(d, desc_add(da, (k >> 2) * kTileBytes + (k & 3) * 32),
                             desc_add(va, (k >> 2) * kTileBytes + (k & 3) * 32), idesc_s);
          ptx::umma_commit(&dp_full[t]);
        }
        __syncwarp();
      };
      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&kv_full[0], 0);
      ptx::tc_fence_after();
      issue_s(t, 0, 0);
      issue_dp(t, 0, 0);
      if (n_sub > 1) issue_s(t, 1, 0);
      int stage = 0;                                   // K/V tile u >> 1
      int nstage = 1 % kStagesB;                       // K/V tile (u >> 1) + 1
      uint32_t nphase = (kStagesB == 1) ? 1 : 0;
      for (int u = 0; u < n_sub; ++u) {
        const int half = u & 1;
        // dP(u+1) as soon as the softmax warps hold dP(u) in registers
        if (u + 1 < n_sub) {
          ptx::mbar_wait(&dp_free[t], u & 1);
          ptx::tc_fence_after();
          issue_dp(t, u + 1, half ? nstage : stage);
        }
        const uint64_t ka = desc_add(dkm0, stage * kT + half * kHalfBytes);
        {
          ptx::mbar_wait(&ds_full[2 * t + half], (u >> 1) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t aa = tmem_base + t * 256 + half * BS;
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {            // dQ[:, 64 hh .. +64) += dS K[:, 64 hh .. +64)
              const uint32_t dd = tmem_base + t * 256 + 192 + hh * 64;
              const uint64_t kh = desc_add(ka, hh * kTileBytes);
              ptx::umma_ts(dd, aa, kh, idesc_q, u > 0 ? 1u : 0u);
#pragma unroll
              for (int k = 1; k < BS / 16; ++k) ptx::umma_ts_acc(dd, aa + k * 8, desc_add(kh, k * 2048), idesc_q);
            }
            if (u + 1 == n_sub) ptx::umma_commit(&dq_final[t]);
          }
          __syncwarp();
          if (u + 2 < n_sub) {
            if (half == 0) {                           // first touch of K/V tile (u >> 1) + 1
              ptx::mbar_wait(&kv_full[nstage], nphase);
              ptx::tc_fence_after();
            }
            issue_s(t, u + 2, nstage);
          }
        }
        if (half == 1 || u + 1 == n_sub) {             // dQ of the K/V tile's last sub-tile has been issued
          if (ptx::elect_one()) ptx::umma_commit(&kv_empty[stage]);
          __syncwarp();
          stage = nstage;
          if (++nstage == kStagesB) { nstage = 0; nphase ^= 1; }
        }
      }
    }
  } else {
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + t * 256;
    static_assert(kThreadsB == 352, "warps 2-9 are the softmax warps, warp 10 the second issuer (NH = 1)");
    const uint32_t dp_addr = lane_addr + 128;
    const uint32_t dq_addr = lane_addr + 192;
    const int row = q0 + t * BQ + r;
    const size_t stat = ((size_t)b * a.heads + h) * a.N + row;
    const float neg_lse2 = (row < a.N) ? -a.lse[stat] * kLog2e : 0.f;
    const float delta = (row < a.N) ? a.delta[stat] : 0.f;
    const float sc = a.scale_log2;
    const uint64_t sc2 = ptx::pack2(sc, sc), nl2 = ptx::pack2(neg_lse2, neg_lse2), nd2 = ptx::pack2(-delta, -delta);
    const uint32_t drop_key = kDrop ? ptx::attn_drop_key(a.drop, bh) : 0u;
    const uint64_t ik2 = ptx::pack2(kDrop ? a.drop.inv_keep : 1.f, kDrop ? a.drop.inv_keep : 1.f);
    for (int u = 0; u < n_sub; ++u) {
      const int bb = u & 1;
      const uint32_t s_addr = lane_addr + bb * BS;
      ptx::mbar_wait(&s_full[2 * t + bb], (u >> 1) & 1);
      ptx::mbar_wait(&dp_full[t], u & 1);
      ptx::tc_fence_after();
      uint32_t s0[32], s1[32], d0[32], d1[32];
      ptx::tmem_ld_32x32(dp_addr, d0);
      ptx::tmem_ld_32x32(dp_addr + 32, d1);
      ptx::tmem_ld_32x32(s_addr, s0);
      ptx::tmem_ld_32x32(s_addr + 32, s1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&dp_free[t]);
      uint32_t pk[16];
      if (kDrop) {        // dS = P (M dP / keep_prob - delta): mask the raw dP bits, the scale rides on the FFMA2 below
        drop_apply_f32(d0, ptx::attn_keep_word(a.drop, drop_key, (uint32_t)row, (uint32_t)(2 * u)));
        drop_apply_f32(d1, ptx::attn_keep_word(a.drop, drop_key, (uint32_t)row, (uint32_t)(2 * u + 1)));
      }
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const uint64_t X = ptx::fma2(ptx::pack2u(s0[i], s0[i + 1]), sc2, nl2);
        const uint64_t Pq = (((i >> 1) % kPolyDq) == kPolyDq - 1) ? ptx::exp2_pair<true>(X) : ptx::exp2_pair<false>(X);
        const uint64_t dPq = ptx::pack2u(d0[i], d0[i + 1]);
        pk[i >> 1] = ptx::pack_bf16x2_pair(ptx::mul2(Pq, kDrop ? ptx::fma2(dPq, ik2, nd2) : ptx::add2(dPq, nd2)));
      }
      ptx::tmem_st_32x16(s_addr, pk);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const uint64_t X = ptx::fma2(ptx::pack2u(s1[i], s1[i + 1]), sc2, nl2);
        const uint64_t Pq = (((i >> 1) % kPolyDq) == kPolyDq - 1) ? ptx::exp2_pair<true>(X) : ptx::exp2_pair<false>(X);
        const uint64_t dPq = ptx::pack2u(d1[i], d1[i + 1]);
        pk[i >> 1] = ptx::pack_bf16x2_pair(ptx::mul2(Pq, kDrop ? ptx::fma2(dPq, ik2, nd2) : ptx::add2(dPq, nd2)));
      }
      ptx::tmem_st_32x16(s_addr + 16, pk);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ds_full[2 * t + bb]);
    }
    ptx::mbar_wait(&dq_final[t], 0);
    ptx::tc_fence_after();
    __nv_bfloat16* op = a.dqkv + ((((size_t)b * a.N + row) * 3 + 0) * a.heads + h) * (NH * 64);
    const float f = a.scale;
#pragma unroll 1
    for (int c = 0; c < NH * 2; ++c) {
      uint32_t o[32];
      ptx::tmem_ld_32x32(dq_addr + c * 32, o);
      ptx::tmem_ld_wait();
      if (row < a.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[i]) * f, __uint_as_float(o[i + 1]) * f);
          w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * f, __uint_as_float(o[i + 3]) * f);
          w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * f, __uint_as_float(o[i + 5]) * f);
          w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * f, __uint_as_float(o[i + 7]) * f);
          *reinterpret_cast<uint4*>(op + c * 32 + i) = w;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

// -(x) as three bf16 terms (hi, mid, lo) in K positions 0..2 of a 16-byte chunk; relative error ~2^-24
__device__ __forceinline__ uint4 neg_split3(float x) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
  return make_uint4(pack_bf16x2(-__bfloat162float(hi), -__bfloat162float(mid)), pack_bf16x2(-__bfloat162float(lo), 0.f), 0u, 0u);
}

// ---------------------------------------------------------------------------------------------- dK, dV
template <int NH, bool kDrop>
__global__ void __launch_bounds__(kThreadsF, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                    const BwdArgs a) {
  constexpr int kStagesB = BwdCfg<NH>::kStages;
  constexpr int kTiles = BwdCfg<NH>::kTiles;
  constexpr uint32_t kT = BwdCfg<NH>::kT;
  constexpr uint32_t kColDK = 128, kColDV = 128 + 64 * NH;   // TMEM columns inside a key tile's block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* sK = smem;                               // kTiles tiles
  uint8_t* sV = sK + kTiles * kT;                   // kTiles tiles
  uint8_t* sQ = sV + kTiles * kT;                   // kStagesB tiles
  uint8_t* sdO = sQ + kStagesB * kT;                // kStagesB tiles
  // Per-query statistics enter through the tensor core as rank-1 updates (one extra K=16 MMA each):
  //   S^T  <- K Q^T  - 1 (lse/scale)^T      dP^T <- V dO^T - 1 delta^T
  // sOnes: [128 x 64] bf16 K-major tile; k-slice 0 has ones in K positions 0..2, k-slice 1 has ones in positions 8..10.
  // sStat: [128 q x 64] tile; k-slice `stage` of row q holds -(3-term bf16 split of lse/scale) in positions 0..2 and
  //        -(split of delta) in positions 8..10, so the same slice serves both updates.
  uint8_t* sOnes = sdO + kStagesB * kT;
  uint8_t* sStat = sOnes + kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + kTileBytes);
  uint64_t* kv_full = bars;                         // 1
  uint64_t* qdo_full = kv_full + 1;                 // kStagesB (2 arrivals: TMA expect_tx + statistics)
  uint64_t* qdo_empty = qdo_full + kStagesB;
  uint64_t* sd_full = qdo_empty + kStagesB;         // 2
  uint64_t* pd_full = sd_full + 2;                  // 2 (128 arrivals)
  uint64_t* dkv_done = pd_full + 2;                 // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dkv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / a.heads, h = bh % a.heads;
  const int k0 = blockIdx.x * kTiles * BKV;
  const int n_sub = a.n_sub;                        // 64-query sub-tiles
  const int n_q = (n_sub + 1) / 2;                  // 128-query TMA tiles

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::prefetch_tmap(&tmap_do);
    ptx::mbar_init(kv_full, 1);
    for (int s = 0; s < kStagesB; ++s) {
      ptx::mbar_init(&qdo_full[s], 2);
      ptx::mbar_init(&qdo_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      ptx::mbar_init(&sd_full[t], 1);
      ptx::mbar_init(&pd_full[t], 256);
      ptx::mbar_init(&dkv_done[t], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + BKV) {
    const int r = threadIdx.x - 64;
    const uint4 ones = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
    uint4* row = reinterpret_cast<uint4*>(sOnes + r * 128);
    row[0 ^ (r & 7)] = ones;  row[1 ^ (r & 7)] = zero;     // k-slice 0: positions 0..2
    row[2 ^ (r & 7)] = zero;  row[3 ^ (r & 7)] = ones;     // k-slice 1: positions 8..10
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns of key tile t (base t*256): S^T [0,64) dP^T [64,128) dK [128, +64 NH) dV [128 + 64 NH, +64 NH)

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(kv_full, 2 * kTiles * kT);
#pragma unroll
      for (int t = 0; t < kTiles; ++t)
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          ptx::tma_load_4d(sK + t * kT + hh * kTileBytes, &tmap_qkv, kv_full, hh * 64, a.heads + h, k0 + t * BKV, b);
          ptx::tma_load_4d(sV + t * kT + hh * kTileBytes, &tmap_qkv, kv_full, hh * 64, 2 * a.heads + h, k0 + t * BKV, b);
        }
    }
    int stage = 0;
    uint32_t phase = 0;
    const float* lse = a.lse + ((size_t)b * a.heads + h) * a.N;
    const float* dlt = a.delta + ((size_t)b * a.heads + h) * a.N;
    const float inv_scale = 1.f / a.scale;
    float xs[4], ys[4];
    auto fetch_stats = [&](int i) {
      float lv[4], dv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = i * BQ + e * 32 + lane;
        const bool ok = row < a.N;
        lv[e] = ok ? __ldg(lse + row) : 0.f;
        dv[e] = ok ? __ldg(dlt + row) : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = i * BQ + e * 32 + lane;
        xs[e] = (row < a.N) ? lv[e] * inv_scale : 1e30f;    // exp2((s - 1e30) * c) = 0 for rows past N
        ys[e] = (row < a.N) ? dv[e] : 0.f;
      }
    };
    fetch_stats(0);
    for (int i = 0; i < n_q; ++i) {
      O2_TL(2 * i, 0, 10);                      // producer: wants the stage of query tile i
      if (lane == 0) {
        ptx::mbar_wait(&qdo_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&qdo_full[stage], 2 * kT);
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          ptx::tma_load_4d(sQ + stage * kT + hh * kTileBytes, &tmap_qkv, &qdo_full[stage], hh * 64, h, i * BQ, b);
          ptx::tma_load_4d(sdO + stage * kT + hh * kTileBytes, &tmap_do, &qdo_full[stage], hh * 64, h, i * BQ, b);
        }
      }
      __syncwarp();
      // the statistics of this tile were fetched one tile ahead (all eight loads in flight together: issued one by one
      // behind the shared-memory stores they cost ~6700 cycles per tile and paced the whole kernel)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = e * 32 + lane;
        uint4* dst = reinterpret_cast<uint4*>(sStat + idx * 128);
        dst[(2 * stage) ^ (idx & 7)] = neg_split3(xs[e]);
        dst[(2 * stage + 1) ^ (idx & 7)] = neg_split3(ys[e]);
        // dropout: dS = P o (dP o M - delta) needs delta outside the accumulator; raw fp32 copy in the unused k-slice 3
        if (kDrop) reinterpret_cast<float*>(dst + (6 ^ (idx & 7)))[stage] = ys[e];
      }
      fetch_stats(i + 1);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&qdo_full[stage]);
      O2_TL(2 * i, 0, 11);                      // producer: TMA issued + statistics written for query tile i
      if (++stage == kStagesB) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    {   // warp-uniform issue loop (see the forward kernel).  ONE issuer for both key tiles: S^T / dP^T are single-
        // buffered here, so each tile's 18 MMAs should run as one burst (two interleaving issuers measured 10% slower)
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BKV, BS, 0, 0);   // S^T / dP^T: N = 64 queries
      const uint32_t idesc_g = ptx::umma_idesc_bf16(BKV, kHD, 0, 1);  // dV / dK: A (TMEM), B = dO / Q (MN-major)
      const uint64_t dk0 = ptx::umma_smem_desc(ptx::smem_u32(sK), 16, 1024);
      const uint64_t dv0 = ptx::umma_smem_desc(ptx::smem_u32(sV), 16, 1024);
      const uint64_t dq0 = ptx::umma_smem_desc(ptx::smem_u32(sQ), 16, 1024);
      const uint64_t ddo0 = ptx::umma_smem_desc(ptx::smem_u32(sdO), 16, 1024);
      const uint64_t dqm0 = ptx::umma_smem_desc(ptx::smem_u32(sQ), 8192, 1024);    // Q as MN-major B
      const uint64_t ddom0 = ptx::umma_smem_desc(ptx::smem_u32(sdO), 8192, 1024);  // dO as MN-major B
      const uint64_t dones = ptx::umma_smem_desc(ptx::smem_u32(sOnes), 16, 1024);
      const uint64_t dstat0 = ptx::umma_smem_desc(ptx::smem_u32(sStat), 16, 1024);
      auto issue_sd = [&](int t, int stage, int half) {
        if (ptx::elect_one()) {
          const uint64_t ka = desc_add(dk0, t * kT), va = desc_add(dv0, t * kT);
          const uint64_t qa = desc_add(dq0, stage * kT + half * kHalfBytes);
          const uint64_t da = desc_add(ddo0, stage * kT + half * kHalfBytes);
          const uint32_t ds_ = tmem_base + t * 256;
          ptx::umma_ss_first(ds_, ka, qa, idesc_s);
#pragma unroll
          for (int k = 1; k < NH * 4; ++k)
            ptx::umma_ss_acc(ds_, desc_add(ka, (k >> 2) * kTileBytes + (k & 3) * 32),
                             desc_add(qa, (k >> 2) * kTileBytes + (k & 3) * 32), idesc_s);
          const uint64_t st = desc_add(dstat0, half * kHalfBytes + stage * 32);
          ptx::umma_ss_acc(ds_, dones, st, idesc_s);                          // - lse / scale
          ptx::umma_ss_first(ds_ + 64, va, da, idesc_s);
#pragma unroll
          for (int k = 1; k < NH * 4; ++k)
            ptx::umma_ss_acc(ds_ + 64, desc_add(va, (k >> 2) * kTileBytes + (k & 3) * 32),
                             desc_add(da, (k >> 2) * kTileBytes + (k & 3) * 32), idesc_s);
          if (!kDrop) ptx::umma_ss_acc(ds_ + 64, desc_add(dones, 32), st, idesc_s);       // - delta
          ptx::umma_commit(&sd_full[t]);
        }
        __syncwarp();
      };
      ptx::mbar_wait(kv_full, 0);
      ptx::mbar_wait(&qdo_full[0], 0);
      ptx::tc_fence_after();
#pragma unroll
      for (int t = 0; t < kTiles; ++t) issue_sd(t, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int u = 0; u < n_sub; ++u) {
        const int half = u & 1;
        const bool more = (u + 1 < n_sub);
        int nstage = stage;
        uint32_t nphase = phase;
        if (half == 1) { if (++nstage == kStagesB) { nstage = 0; nphase ^= 1; } }
        const uint64_t dam = desc_add(ddom0, stage * kT + half * kHalfBytes);
        const uint64_t qam = desc_add(dqm0, stage * kT + half * kHalfBytes);
#pragma unroll
        for (int t = 0; t < kTiles; ++t) {
          O2_TL(u, t, 0);                       // issuer starts waiting for P^T / dS^T of (u, t)
          ptx::mbar_wait(&pd_full[t], u & 1);
          ptx::tc_fence_after();
          O2_TL(u, t, 1);                       // issuer saw pd_full
          if (ptx::elect_one()) {
            const uint32_t base = tmem_base + t * 256;
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {          // dV[:, 64 hh .. +64) += P^T dO[:, 64 hh .. +64)
              const uint32_t dd = base + kColDV + hh * 64;
              const uint64_t bh_ = desc_add(dam, hh * kTileBytes);
              ptx::umma_ts(dd, base, bh_, idesc_g, u > 0 ? 1u : 0u);
#pragma unroll
              for (int k = 1; k < BS / 16; ++k)
                ptx::umma_ts_acc(dd, base + (k >> 1) * 32 + (k & 1) * 8, desc_add(bh_, k * 2048), idesc_g);
            }
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {          // dK[:, 64 hh .. +64) += dS^T Q[:, 64 hh .. +64)
              const uint32_t dd = base + kColDK + hh * 64;
              const uint64_t bh_ = desc_add(qam, hh * kTileBytes);
              ptx::umma_ts(dd, base + 64, bh_, idesc_g, u > 0 ? 1u : 0u);
#pragma unroll
              for (int k = 1; k < BS / 16; ++k)
                ptx::umma_ts_acc(dd, base + 64 + (k >> 1) * 32 + (k & 1) * 8, desc_add(bh_, k * 2048), idesc_g);
            }
            ptx::umma_commit(&dkv_done[t]);
          }
          __syncwarp();
          if (more) {
            O2_TL(u, t, 3);                     // dV / dK issued
            if (t == 0 && half == 1) {                // first touch of the next Q / dO tile
              ptx::mbar_wait(&qdo_full[nstage], nphase);
              ptx::tc_fence_after();
            }
            O2_TL(u, t, 9);                     // next Q / dO tile present
            issue_sd(t, nstage, half ^ 1);
          }
          O2_TL(u, t, 2);                       // all 18 MMAs of (u, t) issued
        }
        if (half == 1 || !more) {
          if (ptx::elect_one()) ptx::umma_commit(&qdo_empty[stage]);
          __syncwarp();
        }
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // All eight softmax warps work on ONE key tile at a time (each thread: one key row, 32 of the 64 query columns), then
    // on the other: the tiles are forced into anti-phase, so the MMA burst of one always overlaps the softmax of the other.
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;            // query columns [32 chalf, +32) of the sub-tile
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sc = a.scale_log2;
    const uint64_t sc2 = ptx::pack2(sc, sc);
    static_assert(!kDrop || BwdCfg<NH>::kStages <= 3, "the raw delta copy lives in k-slice 3 of the statistics tile");
    const uint32_t drop_key = kDrop ? ptx::attn_drop_key(a.drop, bh) : 0u;
    int dstage = 0;                              // stage of the 128-query tile that holds sub-tile u
    for (int u = 0; u < n_sub; ++u) {
#pragma unroll
      for (int t = 0; t < kTiles; ++t) {
        const uint32_t st_addr = lane_addr + t * 256 + chalf * 32;       // this thread's S^T columns; P^T goes over their head
        const uint32_t dp_addr = st_addr + 64;
        if (warp == 2) O2_TL(u, t, 4);          // softmax starts waiting for S^T / dP^T of (u, t)
        ptx::mbar_wait(&sd_full[t], u & 1);
        ptx::tc_fence_after();
        if (warp == 2) O2_TL(u, t, 5);          // saw sd_full
        uint32_t sv_[32], dv_[32];
        ptx::tmem_ld_32x32(st_addr, sv_);
        ptx::tmem_ld_32x32(dp_addr, dv_);
        ptx::tmem_ld_wait();
        if (warp == 2) O2_TL(u, t, 6);          // TMEM loads landed
        uint32_t pk[16], dk[16];
        if (kDrop) {
          // thread = key row: lane l builds the keep word of query qb + l for this warp's 32-key block, every thread
          // reads the words of its 32 queries and tests its own key's bit
          const int key_ = k0 + t * BKV + r;
          const int qb = u * BS + chalf * 32;
          const uint32_t kbit = 1u << ptx::attn_keep_bit((uint32_t)key_);
          // the 32 words of this warp's (32 keys x 32 queries) block go through a private 128-byte shared-memory slot:
          // one store + eight 16-byte broadcast loads instead of 32 shuffles
          uint32_t* wslot = reinterpret_cast<uint32_t*>(smem + (BwdCfg<NH>::kDkvSmem - 1024)) + (warp - 2) * 32;
          __syncwarp();                                 // the previous block's loads of this slot are done
          wslot[lane] = ptx::attn_keep_word(a.drop, drop_key, (uint32_t)(qb + lane), (uint32_t)(key_ >> 5));
          __syncwarp();
          const float* dl = reinterpret_cast<const float*>(sStat) + dstage;
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const uint4 w4 = reinterpret_cast<const uint4*>(wslot)[i4];        // keep words of queries qb + 4 i4 .. + 3
#pragma unroll
            for (int hp = 0; hp < 2; ++hp) {
              const int i = 4 * i4 + 2 * hp;
              const uint64_t X = ptx::mul2(ptx::pack2u(sv_[i], sv_[i + 1]), sc2);
              const uint64_t Pq = ptx::exp2_pair<false>(X);
              const bool k0_ = ((hp ? w4.z : w4.x) & kbit) != 0u;
              const bool k1_ = ((hp ? w4.w : w4.y) & kbit) != 0u;
              float p0, p1;
              ptx::unpack2(Pq, p0, p1);
              const int qi = (u & 1) * BS + chalf * 32 + i;            // row of the 128-query statistics tile
              const float de0 = dl[(qi * 128 + ((6 ^ (qi & 7)) * 16)) >> 2];
              const float de1 = dl[((qi + 1) * 128 + ((6 ^ ((qi + 1) & 7)) * 16)) >> 2];
              const float g0 = k0_ ? __uint_as_float(dv_[i]) * a.drop.inv_keep : 0.f;
              const float g1 = k1_ ? __uint_as_float(dv_[i + 1]) * a.drop.inv_keep : 0.f;
              pk[i >> 1] = pack_bf16x2(k0_ ? p0 : 0.f, k1_ ? p1 : 0.f);
              dk[i >> 1] = pack_bf16x2(p0 * (g0 - de0), p1 * (g1 - de1));
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint64_t X = ptx::mul2(ptx::pack2u(sv_[i], sv_[i + 1]), sc2);
            const uint64_t Pq = (((i >> 1) % kPolyDkv) == kPolyDkv - 1) ? ptx::exp2_pair<true>(X) : ptx::exp2_pair<false>(X);
            pk[i >> 1] = ptx::pack_bf16x2_pair(Pq);
            dk[i >> 1] = ptx::pack_bf16x2_pair(ptx::mul2(Pq, ptx::pack2u(dv_[i], dv_[i + 1])));
          }
        }
        ptx::tmem_st_32x16(st_addr, pk);
        ptx::tmem_st_32x16(dp_addr, dk);
        if (warp == 2) O2_TL(u, t, 7);          // math done, stores issued
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&pd_full[t]);
        if (warp == 2) O2_TL(u, t, 8);          // arrived on pd_full
      }
      if (u & 1) { if (++dstage == kStagesB) dstage = 0; }
    }
    // epilogue.  Two key tiles: warps 2-5 drain tile 0, warps 6-9 tile 1 (dK then dV).  One key tile: warps 2-5 drain
    // dK, warps 6-9 dV.
    const int t = (kTiles == 2) ? chalf : 0;
    const uint32_t st_addr = lane_addr + t * 256;
    ptx::mbar_wait(&dkv_done[t], (n_sub - 1) & 1);
    ptx::tc_fence_after();
    const int key = k0 + t * BKV + r;
    const int w_lo = (kTiles == 2) ? 1 : 1 + chalf, w_hi = (kTiles == 2) ? 2 : 1 + chalf;
#pragma unroll 1
    for (int which = w_lo; which <= w_hi; ++which) {    // 1: dK (scaled), 2: dV
      __nv_bfloat16* op = a.dqkv + ((((size_t)b * a.N + key) * 3 + which) * a.heads + h) * (NH * 64);
      const float f = (which == 1) ? a.scale : (kDrop ? a.drop.inv_keep : 1.f);   // dV = P_kept^T dO / keep_prob
#pragma unroll 1
      for (int c = 0; c < NH * 2; ++c) {
        uint32_t o[32];
        ptx::tmem_ld_32x32(st_addr + (which == 1 ? kColDK : kColDV) + c * 32, o);
        ptx::tmem_ld_wait();
        if (key < a.N) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[i]) * f, __uint_as_float(o[i + 1]) * f);
            w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * f, __uint_as_float(o[i + 3]) * f);
            w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * f, __uint_as_float(o[i + 5]) * f);
            w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * f, __uint_as_float(o[i + 7]) * f);
            *reinterpret_cast<uint4*>(op + c * 32 + i) = w;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------- dK, dV (hd 64, v2)
// Same mathematics as attn_bwd_dkv_kernel, re-organised so that EVERY tensor-core instruction takes its A operand from
// tensor memory.  A CTA owns ONE 128-key tile and keeps K and V themselves (bf16, 32 packed columns each) plus the
// rank-1 "ones" slices in TMEM; S^T / dP^T are double-buffered over consecutive 64-query sub-tiles instead of over two
// key tiles.  An SS-MMA of 128 x 64 x 16 reads 4 KiB of A and 2 KiB of B from shared memory and runs at the
// shared-memory bandwidth (58 cycles instead of 32); here only the 2 KiB of Q / dO are read, so the 18 MMAs of a
// (key tile, query sub-tile) block take 18 x 32 instead of 10 x 58 + 8 x 32 cycles of the tensor pipe.
// TMEM columns: buffer b: S^T [128 b, +64) dP^T [128 b + 64, +64) | dK [256,320) dV [320,384) | K [384,416) V [416,448)
// | ones slice 0 [448,456) slice 1 [456,464).
template <bool kDrop>
__global__ void __launch_bounds__(kThreadsB, 1)
attn_bwd_dkv_tm_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                       const BwdArgs a) {
  constexpr int kStagesB = kDrop ? 3 : 4;     // the dropout variant keeps a raw delta copy in k-slice 3 of the statistics tile
  constexpr uint32_t kT = kTileBytes;
  constexpr uint32_t kColDK = 256, kColDV = 320, kColK = 384, kColV = 416, kColOnes = 448;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                               // kStagesB tiles of 128 queries
  uint8_t* sdO = sQ + kStagesB * kT;                // kStagesB tiles
  uint8_t* sStat = sdO + kStagesB * kT;             // [128 q x 64]: k-slice `stage` = -(split lse/scale | split delta)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + kTileBytes);
  uint64_t* kv_ready = bars;                        // 1 (256 arrivals: K, V and the ones are in TMEM)
  uint64_t* qdo_full = kv_ready + 1;                // kStagesB (2 arrivals: TMA expect_tx + statistics)
  uint64_t* qdo_empty = qdo_full + kStagesB;
  uint64_t* sd_full = qdo_empty + kStagesB;         // 2 (per buffer)
  uint64_t* pd_full = sd_full + 2;                  // 2 (256 arrivals)
  uint64_t* dkv_done = pd_full + 2;                 // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dkv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / a.heads, h = bh % a.heads;
  const int k0 = blockIdx.x * BKV;
  const int n_sub = a.n_sub;                        // 64-query sub-tiles
  const int n_q = (n_sub + 1) / 2;                  // 128-query TMA tiles

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::prefetch_tmap(&tmap_do);
    ptx::mbar_init(kv_ready, 256);
    for (int s = 0; s < kStagesB; ++s) {
      ptx::mbar_init(&qdo_full[s], 2);
      ptx::mbar_init(&qdo_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      ptx::mbar_init(&sd_full[t], 1);
      ptx::mbar_init(&pd_full[t], 256);
      ptx::mbar_init(&dkv_done[t], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (Q / dO tiles of 128 queries)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_q; ++i) {
        ptx::mbar_wait(&qdo_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&qdo_full[stage], 2 * kT);
        ptx::tma_load_4d(sQ + stage * kT, &tmap_qkv, &qdo_full[stage], 0, h, i * BQ, b);
        ptx::tma_load_4d(sdO + stage * kT, &tmap_do, &qdo_full[stage], 0, h, i * BQ, b);
        if (++stage == kStagesB) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------ statistics warp: the second arrival on qdo_full.
    // A 128-query tile lasts ~1.2 us here (one key tile per CTA); TMA issue + split + shared-memory stores + proxy fence
    // in ONE warp took ~1.7 us per tile and paced the whole kernel, so the statistics have their own warp, and they are
    // fetched two tiles ahead into alternating register sets.
    int stage = 0;
    uint32_t phase = 0;
    const float* lse = a.lse + ((size_t)b * a.heads + h) * a.N;
    const float* dlt = a.delta + ((size_t)b * a.heads + h) * a.N;
    const float inv_scale = 1.f / a.scale;
    float xa[4], ya[4], xb[4], yb[4];
    auto fetch_stats = [&](int i, float (&xs)[4], float (&ys)[4]) {
      float lv[4], dv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = i * BQ + e * 32 + lane;
        const bool ok = row < a.N;
        lv[e] = ok ? __ldg(lse + row) : 0.f;
        dv[e] = ok ? __ldg(dlt + row) : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = i * BQ + e * 32 + lane;
        xs[e] = (row < a.N) ? lv[e] * inv_scale : 1e30f;    // exp2((s - 1e30) * c) = 0 for rows past N
        ys[e] = (row < a.N) ? dv[e] : 0.f;
      }
    };
    auto produce = [&](const float (&xs)[4], const float (&ys)[4]) {
      ptx::mbar_wait(&qdo_empty[stage], phase ^ 1);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = e * 32 + lane;
        uint4* dst = reinterpret_cast<uint4*>(sStat + idx * 128);
        dst[(2 * stage) ^ (idx & 7)] = neg_split3(xs[e]);
        dst[(2 * stage + 1) ^ (idx & 7)] = neg_split3(ys[e]);
        if (kDrop) reinterpret_cast<float*>(dst + (6 ^ (idx & 7)))[stage] = ys[e];
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&qdo_full[stage]);
      if (++stage == kStagesB) { stage = 0; phase ^= 1; }
    };
    fetch_stats(0, xa, ya);
    fetch_stats(1, xb, yb);
    for (int i = 0; i < n_q; i += 2) {
      produce(xa, ya);
      fetch_stats(i + 2, xa, ya);
      if (i + 1 < n_q) {
        produce(xb, yb);
        fetch_stats(i + 3, xb, yb);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = ptx::umma_idesc_bf16(BKV, BS, 0, 0);   // S^T / dP^T: N = 64 queries
    const uint32_t idesc_g = ptx::umma_idesc_bf16(BKV, kHD, 0, 1);  // dV / dK: A (TMEM), B = dO / Q (MN-major)
    const uint64_t dq0 = ptx::umma_smem_desc(ptx::smem_u32(sQ), 16, 1024);
    const uint64_t ddo0 = ptx::umma_smem_desc(ptx::smem_u32(sdO), 16, 1024);
    const uint64_t dqm0 = ptx::umma_smem_desc(ptx::smem_u32(sQ), 8192, 1024);    // Q as MN-major B
    const uint64_t ddom0 = ptx::umma_smem_desc(ptx::smem_u32(sdO), 8192, 1024);  // dO as MN-major B
    const uint64_t dstat0 = ptx::umma_smem_desc(ptx::smem_u32(sStat), 16, 1024);
    // S^T / dP^T of the sub-tile that lives in rows [64 half, +64) of Q / dO stage `stage`, into buffer `buf`
    auto issue_sd = [&](int buf, int stage, int half) {
      if (ptx::elect_one()) {
        const uint64_t qa = desc_add(dq0, stage * kT + half * kHalfBytes);
        const uint64_t da = desc_add(ddo0, stage * kT + half * kHalfBytes);
        const uint64_t st = desc_add(dstat0, half * kHalfBytes + stage * 32);
        const uint32_t ds_ = tmem_base + buf * 128;
        ptx::umma_ts(ds_, tmem_base + kColK, qa, idesc_s, 0u);
#pragma unroll
        for (int k = 1; k < 4; ++k) ptx::umma_ts_acc(ds_, tmem_base + kColK + k * 8, desc_add(qa, k * 32), idesc_s);
        ptx::umma_ts_acc(ds_, tmem_base + kColOnes, st, idesc_s);                    // - lse / scale
        ptx::umma_ts(ds_ + 64, tmem_base + kColV, da, idesc_s, 0u);
#pragma unroll
        for (int k = 1; k < 4; ++k) ptx::umma_ts_acc(ds_ + 64, tmem_base + kColV + k * 8, desc_add(da, k * 32), idesc_s);
        if (!kDrop) ptx::umma_ts_acc(ds_ + 64, tmem_base + kColOnes + 8, st, idesc_s);   // - delta
        ptx::umma_commit(&sd_full[buf]);
      }
      __syncwarp();
    };
    ptx::mbar_wait(kv_ready, 0);
    ptx::mbar_wait(&qdo_full[0], 0);
    ptx::tc_fence_after();
    issue_sd(0, 0, 0);
    if (n_sub > 1) issue_sd(1, 0, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int u = 0; u < n_sub; ++u) {
      const int buf = u & 1;                          // = half: sub-tile u is rows [64 buf, +64) of its 128-query tile
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == kStagesB) { nstage = 0; nphase ^= 1; }
      const uint64_t dam = desc_add(ddom0, stage * kT + buf * kHalfBytes);
      const uint64_t qam = desc_add(dqm0, stage * kT + buf * kHalfBytes);
      ptx::mbar_wait(&pd_full[buf], (u >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t base = tmem_base + buf * 128;
        ptx::umma_ts(tmem_base + kColDV, base, dam, idesc_g, u > 0 ? 1u : 0u);           // dV += P^T dO
#pragma unroll
        for (int k = 1; k < BS / 16; ++k)
          ptx::umma_ts_acc(tmem_base + kColDV, base + (k >> 1) * 32 + (k & 1) * 8, desc_add(dam, k * 2048), idesc_g);
        ptx::umma_ts(tmem_base + kColDK, base + 64, qam, idesc_g, u > 0 ? 1u : 0u);       // dK += dS^T Q
#pragma unroll
        for (int k = 1; k < BS / 16; ++k)
          ptx::umma_ts_acc(tmem_base + kColDK, base + 64 + (k >> 1) * 32 + (k & 1) * 8, desc_add(qam, k * 2048), idesc_g);
        ptx::umma_commit(&dkv_done[buf]);
      }
      __syncwarp();
      if (u + 2 < n_sub) {                            // this buffer's next sub-tile (u + 2) lives in the next 128-query tile
        if (buf == 0) {                               // first touch of that tile
          ptx::mbar_wait(&qdo_full[nstage], nphase);
          ptx::tc_fence_after();
        }
        issue_sd(buf, nstage, buf);
      }
      if (buf == 1 || u + 1 == n_sub) {               // both halves of this tile have been consumed
        if (ptx::elect_one()) ptx::umma_commit(&qdo_empty[stage]);
        __syncwarp();
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // eight softmax warps, all on the same sub-tile: thread = one key row, 32 of the 64 query columns
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;            // query columns [32 chalf, +32) of the sub-tile
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sc = a.scale_log2;
    const uint64_t sc2 = ptx::pack2(sc, sc);
    const int key = k0 + r;
    {   // K row (warps 2-5) / V row (warps 6-9) of this thread's key -> its TMEM lane; warps 2-5 also write the ones
      uint32_t w[32];
      if (key < a.N) {
        const uint4* src = reinterpret_cast<const uint4*>(a.qkv + ((((size_t)b * a.N + key) * 3 + 1 + chalf) * a.heads + h) * kHD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 v = src[i];
          w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = 0u;
      }
      ptx::tmem_st_32x32(lane_addr + (chalf == 0 ? kColK : kColV), w);
      if (chalf == 0) {
        // ones in K positions 0..2 of slice 0 and 8..10 of slice 1 (bf16 1.0 = 0x3F80)
        const uint32_t ones[16] = {0x3F803F80u, 0x00003F80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0x3F803F80u, 0x00003F80u, 0u, 0u};
        ptx::tmem_st_32x16(lane_addr + kColOnes, ones);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(kv_ready);
    }
    const uint32_t drop_key = kDrop ? ptx::attn_drop_key(a.drop, bh) : 0u;
    int dstage = 0;                              // stage of the 128-query tile that holds sub-tile u
    for (int u = 0; u < n_sub; ++u) {
      const int buf = u & 1;
      const uint32_t st_addr = lane_addr + buf * 128 + chalf * 32;       // this thread's S^T columns; P^T goes over their head
      const uint32_t dp_addr = st_addr + 64;
      ptx::mbar_wait(&sd_full[buf], (u >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t sv_[32], dv_[32];
      ptx::tmem_ld_32x32(st_addr, sv_);
      ptx::tmem_ld_32x32(dp_addr, dv_);
      ptx::tmem_ld_wait();
      uint32_t pk[16], dk[16];
      if (kDrop) {
        const int qb = u * BS + chalf * 32;
        const uint32_t kbit = 1u << ptx::attn_keep_bit((uint32_t)key);
        const uint32_t my_word = ptx::attn_keep_word(a.drop, drop_key, (uint32_t)(qb + lane), (uint32_t)(key >> 5));
        const float* dl = reinterpret_cast<const float*>(sStat) + dstage;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t X = ptx::mul2(ptx::pack2u(sv_[i], sv_[i + 1]), sc2);
          const uint64_t Pq = ptx::exp2_pair<false>(X);
          const bool k0_ = (__shfl_sync(0xffffffffu, my_word, i) & kbit) != 0u;
          const bool k1_ = (__shfl_sync(0xffffffffu, my_word, i + 1) & kbit) != 0u;
          float p0, p1;
          ptx::unpack2(Pq, p0, p1);
          const int qi = (u & 1) * BS + chalf * 32 + i;            // row of the 128-query statistics tile
          const float de0 = dl[(qi * 128 + ((6 ^ (qi & 7)) * 16)) >> 2];
          const float de1 = dl[((qi + 1) * 128 + ((6 ^ ((qi + 1) & 7)) * 16)) >> 2];
          const float g0 = k0_ ? __uint_as_float(dv_[i]) * a.drop.inv_keep : 0.f;
          const float g1 = k1_ ? __uint_as_float(dv_[i + 1]) * a.drop.inv_keep : 0.f;
          pk[i >> 1] = pack_bf16x2(k0_ ? p0 : 0.f, k1_ ? p1 : 0.f);
          dk[i >> 1] = pack_bf16x2(p0 * (g0 - de0), p1 * (g1 - de1));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t X = ptx::mul2(ptx::pack2u(sv_[i], sv_[i + 1]), sc2);
          const uint64_t Pq = (((i >> 1) % kPolyDkv) == kPolyDkv - 1) ? ptx::exp2_pair<true>(X) : ptx::exp2_pair<false>(X);
          pk[i >> 1] = ptx::pack_bf16x2_pair(Pq);
          dk[i >> 1] = ptx::pack_bf16x2_pair(ptx::mul2(Pq, ptx::pack2u(dv_[i], dv_[i + 1])));
        }
      }
      ptx::tmem_st_32x16(st_addr, pk);
      ptx::tmem_st_32x16(dp_addr, dk);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&pd_full[buf]);
      if (u & 1) { if (++dstage == kStagesB) dstage = 0; }
    }
    // epilogue: warps 2-5 drain dK (scaled), warps 6-9 dV
    ptx::mbar_wait(&dkv_done[(n_sub - 1) & 1], ((n_sub - 1) >> 1) & 1);
    ptx::tc_fence_after();
    const int which = 1 + chalf;
    __nv_bfloat16* op = a.dqkv + ((((size_t)b * a.N + key) * 3 + which) * a.heads + h) * kHD;
    const float f = (which == 1) ? a.scale : (kDrop ? a.drop.inv_keep : 1.f);   // dV = P_kept^T dO / keep_prob
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      ptx::tmem_ld_32x32(lane_addr + (which == 1 ? kColDK : kColDV) + c * 32, o);
      ptx::tmem_ld_wait();
      if (key < a.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[i]) * f, __uint_as_float(o[i + 1]) * f);
          w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * f, __uint_as_float(o[i + 3]) * f);
          w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * f, __uint_as_float(o[i + 5]) * f);
          w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * f, __uint_as_float(o[i + 7]) * f);
          *reinterpret_cast<uint4*>(op + c * 32 + i) = w;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}
constexpr uint32_t kDkvTmSmem = (2 * 4 + 1) * kTileBytes + 1024 + 256;

int make_qkv_tmap(CUtensorMap* tm, const void* qkv, int B, int N, int heads, int hd, int box_rows) {
  uint64_t dims[4] = {(uint64_t)hd, (uint64_t)(3 * heads), (uint64_t)N, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)hd * 2, (uint64_t)3 * heads * hd * 2, (uint64_t)N * 3 * heads * hd * 2};
  uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};       // 64-column half-tiles (one 128-byte swizzle atom per row)
  return o2_make_tmap(tm, qkv, 2, 4, dims, str, box, 1);
}



#ifdef O2_TIMELINE
extern "C" int o2_debug_timeline(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, g_timeline, sizeof(long long) * n);
}
#endif

template <int NH, bool kDrop>
int launch_fwd(const CUtensorMap& tm, const FwdArgs& a, dim3 grid, cudaStream_t st) {
  O2_SET_SMEM_ONCE((attn_fwd_tc_kernel<NH, kDrop>), FwdCfg<NH>::kSmemBytes);
  attn_fwd_tc_kernel<NH, kDrop><<<grid, kThreadsB, FwdCfg<NH>::kSmemBytes, st>>>(tm, a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

// p -> (site key, 16-bit drop threshold as two byte comparators + dither, exact keep scale); p == 0 -> thr16 = 0 (kernels without the kDrop code)
ptx::AttnDrop make_drop(float p, uint64_t seed, uint32_t site, int N) { return ptx::make_attn_drop(p, seed, site, N); }

template <int NH, bool kDrop>
int launch_bwd(const CUtensorMap& tm_qkv, const CUtensorMap& tm_do, const BwdArgs& a, int parts, cudaStream_t st) {
  using Cfg = BwdCfg<NH>;
  O2_SET_SMEM_ONCE((attn_bwd_dq_kernel<NH, kDrop>), Cfg::kDqSmem);
  O2_SET_SMEM_ONCE((attn_bwd_dkv_kernel<NH, kDrop>), Cfg::kDkvSmem + (kDrop ? 1024 : 0));   // + the keep-word slots of the 8 softmax warps
  const int rows_per_cta = Cfg::kTiles * BQ;
  dim3 grid((a.N + rows_per_cta - 1) / rows_per_cta, a.B * a.heads);
  if (parts & O2_ATTN_BWD_DKV) {
    // the all-TS kernel (K / V in TMEM, one key tile per CTA) measures the same as the two-tile kernel in isolation
    // (1017 vs 1024 TFLOP/s) and inside the step (A/B on one box: 423.0 / 424.1 vs 424.2 / 422.3 ms) while doubling the
    // L2 -> SM traffic for Q / dO, so it stays opt-in: both are bound by the 2-buffer softmax <-> MMA hand-off chain,
    // not by the tensor pipe (54 % busy) or MUFU (51 %).  profiles/r01_attn_experiments.md
    static const bool use_tm = getenv("O2_DKV_TM") != nullptr;
    if (NH == 1 && use_tm) {
      O2_SET_SMEM_ONCE((attn_bwd_dkv_tm_kernel<kDrop>), kDkvTmSmem);
      dim3 grid2((a.N + BKV - 1) / BKV, a.B * a.heads);
      attn_bwd_dkv_tm_kernel<kDrop><<<grid2, kThreadsB, kDkvTmSmem, st>>>(tm_qkv, tm_do, a);
    } else {
      attn_bwd_dkv_kernel<NH, kDrop><<<grid, kThreadsF, Cfg::kDkvSmem + (kDrop ? 1024 : 0), st>>>(tm_qkv, tm_do, a);
    }
    O2_LAUNCH_CHECK();
  }
  if (parts & O2_ATTN_BWD_DQ) {
    attn_bwd_dq_kernel<NH, kDrop><<<grid, Cfg::kDqThreads, Cfg::kDqSmem, st>>>(tm_qkv, tm_do, a);
    O2_LAUNCH_CHECK();
  }
  return O2_OK;
}

#include "attn_bwd_fused.cuh"

// attn_bwd_v3.cuh (N = 128 score MMAs, statistics from shared memory) is an experiment that did NOT win: 942 vs 958 TFLOP/s
// at B = 2, 854 vs 895 at B = 8 (profiles/r02_attn_experiments.md).  It is compiled only into A/B builds
// (tools/build_variant.sh v3 -DO2_ATTN_BWD_V3) and selected there with O2_ATTN_BWD_FUSED_V3=1.
#ifdef O2_ATTN_BWD_V3
#include "attn_bwd_v3.cuh"
#endif

template <bool kDrop>
int launch_bwd_fused(const CUtensorMap& tm_qkv, const CUtensorMap& tm_do, const CUtensorMap& tm_dq, const BwdArgs& a,
                     cudaStream_t st) {
  dim3 grid((a.N + BKV - 1) / BKV, a.B * a.heads);
#ifdef O2_ATTN_BWD_V3
  static const bool use_v3 = getenv("O2_ATTN_BWD_FUSED_V3") != nullptr;
  if (use_v3) {
    O2_SET_SMEM_ONCE((attn_bwd_v3_kernel<kDrop>), kV3Smem);
    attn_bwd_v3_kernel<kDrop><<<grid, kV3Threads, kV3Smem, st>>>(tm_qkv, tm_do, tm_dq, a);
    O2_LAUNCH_CHECK();
    return O2_OK;
  }
#endif
  O2_SET_SMEM_ONCE((attn_bwd_fused_kernel<kDrop>), kFusedSmem);
  attn_bwd_fused_kernel<kDrop><<<grid, kFusedThreads, kFusedSmem, st>>>(tm_qkv, tm_do, tm_dq, a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

}  // namespace

// head dim 256 (interm_10b): attn_tc256.cu
int o2_attn_fwd_tc256(const void* qkv, void* out, float* lse, int B, int N, int heads, float scale, cudaStream_t st);
int o2_attn_bwd_tc256(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta, int B,
                      int N, int heads, float scale, int parts, cudaStream_t st);

int o2_attn_fwd_tc(const void* qkv, void* out, float* lse, int B, int N, int heads, int hd, float scale, float p_drop,
                   uint64_t seed, uint32_t site, cudaStream_t st) {
  if (hd == 256) {
    if (p_drop > 0.f) O2_FAIL(O2_ERR_UNSUPPORTED, "attn_fwd_tc: head dim 256 has no dropout variant (use the fp32 arm)");
    return o2_attn_fwd_tc256(qkv, out, lse, B, N, heads, scale, st);
  }
  O2_REQUIRE(hd == 64 || hd == 128, "attn_fwd_tc: head dim %d not supported (64, 128 or 256)", hd);
  O2_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "attn_fwd_tc: pointers must be 16-byte aligned");
  O2_REQUIRE((long long)B * heads <= 65535, "attn_fwd_tc: B*heads too large");
  CUtensorMap tm;
  int rc = make_qkv_tmap(&tm, qkv, B, N, heads, hd, BQ);
  if (rc) return rc;
  FwdArgs a;
  a.qkv = (const __nv_bfloat16*)qkv;
  a.out = (__nv_bfloat16*)out; a.lse = lse; a.B = B; a.N = N; a.heads = heads;
  a.n_sub = (N + BS - 1) / BS;
  a.scale_log2 = scale * kLog2e;
  dim3 grid((N + 2 * BQ - 1) / (2 * BQ), B * heads);
  a.drop = make_drop(p_drop, seed, site, N);
  O2_REQUIRE((long long)N * a.drop.nkb < (1ll << 32), "attn_fwd_tc: N=%d too large for the dropout word index", N);
  if (a.drop.thr16 > 0) return hd == 64 ? launch_fwd<1, true>(tm, a, grid, st) : launch_fwd<2, true>(tm, a, grid, st);
  return hd == 64 ? launch_fwd<1, false>(tm, a, grid, st) : launch_fwd<2, false>(tm, a, grid, st);
}

int o2_attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta, int B,
                   int N, int heads, int hd, float scale, int parts, float p_drop, uint64_t seed, uint32_t site,
                   cudaStream_t st) {
  if (hd == 256) {
    if (p_drop > 0.f) O2_FAIL(O2_ERR_UNSUPPORTED, "attn_bwd_tc: head dim 256 has no dropout variant (use the fp32 arm)");
    return o2_attn_bwd_tc256(qkv, out, dout, lse, dqkv, delta, B, N, heads, scale, parts, st);
  }
  O2_REQUIRE(hd == 64 || hd == 128, "attn_bwd_tc: head dim %d not supported (64, 128 or 256)", hd);
  O2_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dout % 16) == 0 &&
                 ((uintptr_t)dqkv % 16) == 0,
             "attn_bwd_tc: pointers must be 16-byte aligned");
  O2_REQUIRE((long long)B * heads <= 65535, "attn_bwd_tc: B*heads too large");
  CUtensorMap tm_qkv, tm_do;
  int rc = make_qkv_tmap(&tm_qkv, qkv, B, N, heads, hd, BQ);
  if (rc) return rc;
  {
    uint64_t dims[4] = {(uint64_t)hd, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)hd * 2, (uint64_t)heads * hd * 2, (uint64_t)N * heads * hd * 2};
    uint32_t box[4] = {64, 1, (uint32_t)BQ, 1};            // 64-column half-tiles, like q / k / v
    rc = o2_make_tmap(&tm_do, dout, 2, 4, dims, str, box, 1);
    if (rc) return rc;
  }
  const long long rows = (long long)B * N * heads;
  if (parts & O2_ATTN_BWD_DELTA) {
    if (hd == 64)
      attn_delta_bf16_kernel<1><<<(unsigned)((rows * 8 + 255) / 256), 256, 0, st>>>(
          (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, delta, B, N, heads);
    else
      attn_delta_bf16_kernel<2><<<(unsigned)((rows * 16 + 255) / 256), 256, 0, st>>>(
          (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, delta, B, N, heads);
    O2_LAUNCH_CHECK();
  }
  BwdArgs a;
  a.qkv = (const __nv_bfloat16*)qkv;
  a.lse = lse; a.delta = delta; a.dqkv = (__nv_bfloat16*)dqkv; a.B = B; a.N = N; a.heads = heads;
  a.n_sub = (N + BS - 1) / BS;
  a.scale = scale; a.scale_log2 = scale * kLog2e;
  a.drop = make_drop(p_drop, seed, site, N);
  O2_REQUIRE((long long)N * a.drop.nkb < (1ll << 32), "attn_bwd_tc: N=%d too large for the dropout word index", N);
  if (a.drop.thr16 > 0)
    return hd == 64 ? launch_bwd<1, true>(tm_qkv, tm_do, a, parts, st) : launch_bwd<2, true>(tm_qkv, tm_do, a, parts, st);
  return hd == 64 ? launch_bwd<1, false>(tm_qkv, tm_do, a, parts, st) : launch_bwd<2, false>(tm_qkv, tm_do, a, parts, st);
}

// One-pass backward (head dim 64): delta, then the fused dK / dV / dQ kernel (dQ partials reduced into dq_accum by TMA),
// then dQ = scale * dq_accum.  parts: O2_ATTN_BWD_DELTA | O2_ATTN_BWD_FUSED | O2_ATTN_BWD_DQ_FINISH.
int o2_attn_bwd_fused_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta,
                         float* dq_accum, int B, int N, int heads, int hd, float scale, int parts, float p_drop,
                         uint64_t seed, uint32_t site, cudaStream_t st) {
  O2_REQUIRE(hd == 64, "attn_bwd_fused: head dim %d not supported (64)", hd);
  O2_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dout % 16) == 0 &&
                 ((uintptr_t)dqkv % 16) == 0 && ((uintptr_t)dq_accum % 16) == 0,
             "attn_bwd_fused: pointers must be 16-byte aligned");
  O2_REQUIRE((long long)B * heads <= 65535, "attn_bwd_fused: B*heads too large");
  const long long rows = (long long)B * N * heads;
  if (parts & O2_ATTN_BWD_DELTA) {
    attn_delta_bf16_kernel<1><<<(unsigned)((rows * 8 + 255) / 256), 256, 0, st>>>(
        (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, delta, B, N, heads);
    O2_LAUNCH_CHECK();
  }
  if (parts & O2_ATTN_BWD_FUSED) {
    CUtensorMap tm_qkv, tm_do, tm_dq;
    int rc = make_qkv_tmap(&tm_qkv, qkv, B, N, heads, hd, BQ);
    if (rc) return rc;
    {
      uint64_t dims[4] = {(uint64_t)hd, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
      uint64_t str[3] = {(uint64_t)hd * 2, (uint64_t)heads * hd * 2, (uint64_t)N * heads * hd * 2};
      uint32_t box[4] = {64, 1, (uint32_t)BQ, 1};
      rc = o2_make_tmap(&tm_do, dout, 2, 4, dims, str, box, 1);
      if (rc) return rc;
    }
    {   // dq_accum [B * heads, N, 64] fp32; box = 32 queries x 32 columns (128-byte rows, 128B swizzle) per drain warp
      uint64_t dims[3] = {64, (uint64_t)N, (uint64_t)B * heads};
      uint64_t str[2] = {64 * 4, (uint64_t)N * 64 * 4};
      uint32_t box[3] = {32, 32, 1};
      rc = o2_make_tmap(&tm_dq, dq_accum, 4, 3, dims, str, box, 1);
      if (rc) return rc;
    }
    O2_CUDA(cudaMemsetAsync(dq_accum, 0, (size_t)rows * 64 * sizeof(float), st));
    BwdArgs a;
    a.qkv = (const __nv_bfloat16*)qkv;
    a.lse = lse; a.delta = delta; a.dqkv = (__nv_bfloat16*)dqkv; a.B = B; a.N = N; a.heads = heads;
    a.n_sub = (N + BS - 1) / BS;
    a.scale = scale; a.scale_log2 = scale * kLog2e;
    a.drop = make_drop(p_drop, seed, site, N);
    O2_REQUIRE((long long)N * a.drop.nkb < (1ll << 32), "attn_bwd_fused: N=%d too large for the dropout word index", N);
    rc = a.drop.thr16 > 0 ? launch_bwd_fused<true>(tm_qkv, tm_do, tm_dq, a, st) : launch_bwd_fused<false>(tm_qkv, tm_do, tm_dq, a, st);
    if (rc) return rc;
  }
  if (parts & O2_ATTN_BWD_DQ_FINISH) {
    attn_dq_finish_kernel<<<(unsigned)((rows * 8 + 255) / 256), 256, 0, st>>>(dq_accum, (__nv_bfloat16*)dqkv, B, N, heads, scale);
    O2_LAUNCH_CHECK();
  }
  return O2_OK;
}
