// One-pass flash-attention backward for head dim 64, version 3 (bf16 operands, fp32 accumulation) -- textually included by
// attn_tc.cu inside its anonymous namespace after attn_bwd_fused.cuh (shares BwdArgs, desc_add and the dropout helpers).
//
// reference semantics: components/attention.py:54-78 (autograd backward of softmax(q k^T hd^-0.5) v, attn_drop included).
//
// NOTE (measured after this was written, tools/mma_rate.cu): an M = 128, K = 16 MMA takes 45.5 cycles at N = 64 (TS; 48.1 SS)
// and 64.1 at N = 128 -- not the 58 / 64 vs 32 assumed below -- so the N = 64 score MMAs of the shipped kernel already ran at
// 70 % of the math rate and the tensor pipe was never the limiter; this version measured 942 vs 958 TFLOP/s and is not shipped.
// Why a third version.  ncu of attn_bwd_fused_kernel (profiles/r02_attn_fused_ncu.md): tensor pipe 52 % busy although the
// issue queue never runs dry.  A tcgen05.mma of 128 x 64 x 16 does 32 cycles of math but takes ~64 (TS: the 4 KiB A slice
// is read from tensor memory at 64 B/clk) or ~58 (SS: 6 KiB from shared memory at 128 B/clk); only N = 128 instructions
// run at the math rate.  The fused kernel works on 64-query sub-tiles, so all of its 44 MMAs per (128 keys x 128 queries)
// block are N = 64: 2768 issue cycles for 1408 cycles of math.  Here the two score GEMMs are N = 128 instructions:
//     S^T  [128 k x 128 q] = K Q^T          4 SS-MMAs                      dP^T = V dO^T   likewise
//     dV  += P^T dO,  dK += dS^T Q         8 TS-MMAs (N = 64) each        dQ   = dS K     8 SS-MMAs (N = 64)
// and the rank-1 statistics MMAs are gone: the softmax threads read -lse / -delta of their query columns from shared
// memory (warp-wide broadcast loads).
// TMEM (all 512 columns): S^T [0,128) | dP^T [128,256) | dK [256,320) | dV [320,384) | dQ [384,448) | P^T [448,512).
// The bf16 P^T has its OWN 64 columns, so S^T is free again as soon as the softmax threads hold it in registers
// (s_loaded): S^T(t+2) is computed while tile t+1 is still in its softmax, and the exponentials of a tile never wait for
// the tensor pipe.  dS^T is written over dP^T (chunk j over the first 16 columns of chunk j: own columns, no hazard).
//     tensor pipe, per query tile t:   dK(t) | dP^T(t+1) | dV(t) | dQ(t) | S^T(t+2)
//     softmax threads, tile t+1:       exp2 (S^T ready)  ....  dS^T (after dP^T(t+1)) | P^T store (after dV(t))
// Shared memory (212 KiB): K 16 | V 16 | Q ring 3 x 16 | dO ring 3 x 16 | dS^T 2 sets x 2 x 16 (A operand of dQ = dS K in
// MN-major form; double-buffered because dQ(t) reads set t & 1 while the softmax threads write tile t+1) | dQ staging 16
// | statistics 3 x 1 KiB.
// dQ partials go to the fp32 accumulator by TMA reduce exactly as in attn_bwd_fused_kernel (same workspace, same finish
// kernel, same reproducibility statement); dK / dV are exact single-owner sums.
// (Sixteen softmax warps with setmaxnreg 64 / 88 were tried: 689 vs 786 TFLOP/s for eight -- the exponentials are MUFU
// bound at 1024 cycles per block whatever the warp count; profiles/r02_attn_experiments.md.)
// warps (480 threads): 0 TMA producer | 1 tcgen05 issuer | 2-9 softmax (+ dK / dV epilogue) | 10 statistics | 11-14 dQ drain
constexpr int kV3Threads = 480;
constexpr int kV3Stages = 3;
constexpr uint32_t kV3Smem = (2 + 2 * kV3Stages + 4 + 1) * kTileBytes + kV3Stages * 1024 + 256 + 1024;
#ifndef O2_POLY_V3
#define O2_POLY_V3 (1 << 20)
#endif
constexpr int kPolyV3 = O2_POLY_V3;

template <bool kDrop>
__global__ void __launch_bounds__(kV3Threads, 1)
attn_bwd_v3_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                   const __grid_constant__ CUtensorMap tmap_dq, const BwdArgs a) {
  constexpr int kSt = kV3Stages;
  constexpr uint32_t kT = kTileBytes;
  constexpr uint32_t kColDP = 128, kColDK = 256, kColDV = 320, kColDQ = 384, kColPT = 448;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;                               // [128 k x 64 d]: A of S^T (K-major) and B of dQ (MN-major)
  uint8_t* sV = sK + kT;                            // [128 k x 64 d]: A of dP^T
  uint8_t* sQ = sV + kT;                            // kSt tiles of 128 queries
  uint8_t* sdO = sQ + kSt * kT;                     // kSt tiles
  uint8_t* sdS = sdO + kSt * kT;                    // 2 sets x 2 atoms [128 k x 64 q] bf16: dS^T rows
  uint8_t* sStage = sdS + 4 * kT;                   // 4 drain warps x [32 q x 32 d] fp32, 128B-swizzled
  float* sStat = reinterpret_cast<float*>(sStage + kT);   // [stage][0..127] = -lse log2e, [128..255] = -delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + kSt * 1024);
  uint64_t* kv_full = bars;                         // 1
  uint64_t* qdo_full = kv_full + 1;                 // kSt (2 arrivals: TMA expect_tx + statistics)
  uint64_t* qdo_empty = qdo_full + kSt;             // kSt
  uint64_t* s_full = qdo_empty + kSt;               // 1
  uint64_t* dp_full = s_full + 1;                   // 1
  uint64_t* pd_full = dp_full + 1;                  // 1 (256 arrivals): P^T and dS^T of the tile are in place
  uint64_t* dq_full = pd_full + 1;                  // 1
  uint64_t* dq_free = dq_full + 1;                  // 1 (128 arrivals): the drain warps hold the partial in registers
  uint64_t* ds_free = dq_free + 1;                  // 2: dQ(t) has read dS^T set t & 1
  uint64_t* dkv_done = ds_free + 2;                 // 1
  uint64_t* s_loaded = dkv_done + 1;                // 1 (256 arrivals): S^T of the tile is in registers
  uint64_t* pt_free = s_loaded + 1;                 // 1: dV(t) has read P^T(t)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pt_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / a.heads, h = bh % a.heads;
  const int kt = blockIdx.x;
  const int k0 = kt * BKV;
  const int n_pairs = (a.N + BQ - 1) / BQ;          // 128-query tiles
  const int p0 = kt % n_pairs;                      // CTAs of one head start at different query tiles (spreads the reduces)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::prefetch_tmap(&tmap_do);
    ptx::prefetch_tmap(&tmap_dq);
    ptx::mbar_init(kv_full, 1);
    for (int s = 0; s < kSt; ++s) {
      ptx::mbar_init(&qdo_full[s], 2);
      ptx::mbar_init(&qdo_empty[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(dp_full, 1);
    ptx::mbar_init(pd_full, 256);
    ptx::mbar_init(dq_full, 1);
    ptx::mbar_init(dq_free, 128);
    ptx::mbar_init(&ds_free[0], 1);
    ptx::mbar_init(&ds_free[1], 1);
    ptx::mbar_init(dkv_done, 1);
    ptx::mbar_init(s_loaded, 256);
    ptx::mbar_init(pt_free, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::mbar_expect_tx(kv_full, 2 * kT);
      ptx::tma_load_4d(sK, &tmap_qkv, kv_full, 0, a.heads + h, k0, b);
      ptx::tma_load_4d(sV, &tmap_qkv, kv_full, 0, 2 * a.heads + h, k0, b);
      int stage = 0;
      uint32_t phase = 0;
      int pp = p0;
      for (int i = 0; i < n_pairs; ++i) {
        ptx::mbar_wait(&qdo_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&qdo_full[stage], 2 * kT);
        ptx::tma_load_4d(sQ + stage * kT, &tmap_qkv, &qdo_full[stage], 0, h, pp * BQ, b);
        ptx::tma_load_4d(sdO + stage * kT, &tmap_do, &qdo_full[stage], 0, h, pp * BQ, b);
        if (++stage == kSt) { stage = 0; phase ^= 1; }
        if (++pp == n_pairs) pp = 0;
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------ statistics warp: the second arrival on qdo_full
    int stage = 0;
    uint32_t phase = 0;
    int pp = p0;
    const float* lse = a.lse + ((size_t)b * a.heads + h) * a.N;
    const float* dlt = a.delta + ((size_t)b * a.heads + h) * a.N;
    for (int i = 0; i < n_pairs; ++i) {
      float lv[4], dv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = pp * BQ + e * 32 + lane;
        lv[e] = row < a.N ? -__ldg(lse + row) * kLog2e : -1e30f;        // 2^(s c - 1e30) = 0 for rows past N
        dv[e] = row < a.N ? -__ldg(dlt + row) : 0.f;
      }
      ptx::mbar_wait(&qdo_empty[stage], phase ^ 1);
      float* st = sStat + stage * 256;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        st[e * 32 + lane] = lv[e];
        st[128 + e * 32 + lane] = dv[e];
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&qdo_full[stage]);
      if (++stage == kSt) { stage = 0; phase ^= 1; }
      if (++pp == n_pairs) pp = 0;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ tcgen05 issuer (warp-uniform loop, one elected lane)
    const uint32_t idesc_s = ptx::umma_idesc_bf16(BKV, BQ, 0, 0);   // S^T / dP^T: N = 128 queries
    const uint32_t idesc_g = ptx::umma_idesc_bf16(BKV, kHD, 0, 1);  // dV / dK: A (TMEM), B = dO / Q (MN-major)
    const uint32_t idesc_q = ptx::umma_idesc_bf16(BQ, kHD, 1, 1);   // dQ: A = dS (MN-major: rows = keys), B = K (MN-major)
    const uint64_t dk0 = ptx::umma_smem_desc(ptx::smem_u32(sK), 16, 1024);
    const uint64_t dv0 = ptx::umma_smem_desc(ptx::smem_u32(sV), 16, 1024);
    const uint64_t dq0 = ptx::umma_smem_desc(ptx::smem_u32(sQ), 16, 1024);
    const uint64_t ddo0 = ptx::umma_smem_desc(ptx::smem_u32(sdO), 16, 1024);
    const uint64_t dqm0 = ptx::umma_smem_desc(ptx::smem_u32(sQ), 8192, 1024);    // Q as MN-major B
    const uint64_t ddom0 = ptx::umma_smem_desc(ptx::smem_u32(sdO), 8192, 1024);  // dO as MN-major B
    const uint64_t dkm0 = ptx::umma_smem_desc(ptx::smem_u32(sK), 8192, 1024);    // K as MN-major B
    // dS: M = 128 queries = two 64-query atoms kT apart, K = keys in 8-row groups of 1024 bytes
    const uint64_t dsd0 = ptx::umma_smem_desc(ptx::smem_u32(sdS), kT, 1024);
    auto issue_s = [&](int stage) {                  // S^T = K Q^T
      if (ptx::elect_one()) {
        const uint64_t qa = desc_add(dq0, stage * kT);
        ptx::umma_ss_first(tmem_base, dk0, qa, idesc_s);
#pragma unroll
        for (int k = 1; k < 4; ++k) ptx::umma_ss_acc(tmem_base, desc_add(dk0, k * 32), desc_add(qa, k * 32), idesc_s);
        ptx::umma_commit(s_full);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int stage) {                 // dP^T = V dO^T
      if (ptx::elect_one()) {
        const uint64_t da = desc_add(ddo0, stage * kT);
        ptx::umma_ss_first(tmem_base + kColDP, dv0, da, idesc_s);
#pragma unroll
        for (int k = 1; k < 4; ++k) ptx::umma_ss_acc(tmem_base + kColDP, desc_add(dv0, k * 32), desc_add(da, k * 32), idesc_s);
        ptx::umma_commit(dp_full);
      }
      __syncwarp();
    };
    ptx::mbar_wait(kv_full, 0);
    ptx::mbar_wait(&qdo_full[0], 0);
    ptx::tc_fence_after();
    issue_s(0);
    issue_dp(0);
    if (n_pairs > 1) {
      ptx::mbar_wait(&qdo_full[1], 0);
      ptx::mbar_wait(s_loaded, 0);                   // S^T(0) is in the softmax threads' registers
      ptx::tc_fence_after();
      issue_s(1);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < n_pairs; ++t) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == kSt) { nstage = 0; nphase ^= 1; }
      int n2stage = nstage + 1;
      uint32_t n2phase = nphase;
      if (n2stage == kSt) { n2stage = 0; n2phase ^= 1; }
      const bool more = t + 1 < n_pairs;
      ptx::mbar_wait(pd_full, t & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {                        // dK += dS^T Q: bf16 dS^T of 32-query chunk j at dP^T columns [32 j, 32 j + 16)
        const uint64_t qam = desc_add(dqm0, stage * kT);
        ptx::umma_ts(tmem_base + kColDK, tmem_base + kColDP, qam, idesc_g, t > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 1; k < BQ / 16; ++k)
          ptx::umma_ts_acc(tmem_base + kColDK, tmem_base + kColDP + (k >> 1) * 32 + (k & 1) * 8, desc_add(qam, k * 2048), idesc_g);
      }
      __syncwarp();
      if (more) issue_dp(nstage);                    // (qdo_full[nstage] was waited for when S^T(t+1) was issued)
      if (ptx::elect_one()) {                        // dV += P^T dO: P^T is K-major in its own 64 columns
        const uint64_t dam = desc_add(ddom0, stage * kT);
        ptx::umma_ts(tmem_base + kColDV, tmem_base + kColPT, dam, idesc_g, t > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 1; k < BQ / 16; ++k)
          ptx::umma_ts_acc(tmem_base + kColDV, tmem_base + kColPT + k * 8, desc_add(dam, k * 2048), idesc_g);
        ptx::umma_commit(pt_free);
        ptx::umma_commit(&qdo_empty[stage]);         // every MMA that reads this Q / dO stage has been issued
      }
      __syncwarp();
      if (t > 0) {                                   // the drain warps hold the previous partial in registers
        ptx::mbar_wait(dq_free, (t - 1) & 1);
        ptx::tc_fence_after();
      }
      if (ptx::elect_one()) {                        // dQ[128 q x 64 d] = dS[128 q x 128 k] K[128 k x 64 d]
#ifndef O2_V3_NODQ                                    // (timing experiment: no dQ MMAs)
        const uint64_t a0 = desc_add(dsd0, (t & 1) * 2 * kT);
        ptx::umma_ss_first(tmem_base + kColDQ, a0, dkm0, idesc_q);
#pragma unroll
        for (int k = 1; k < BKV / 16; ++k)
          ptx::umma_ss_acc(tmem_base + kColDQ, desc_add(a0, k * 2048), desc_add(dkm0, k * 2048), idesc_q);
#endif
        ptx::umma_commit(dq_full);
        ptx::umma_commit(&ds_free[t & 1]);
        if (!more) ptx::umma_commit(dkv_done);
      }
      __syncwarp();
      if (t + 2 < n_pairs) {                         // S^T(t+2): its columns are free once tile t+1 is in registers
        ptx::mbar_wait(&qdo_full[n2stage], n2phase);
        ptx::mbar_wait(s_loaded, (t + 1) & 1);
        ptx::tc_fence_after();
        issue_s(n2stage);
      }
      stage = nstage;
      phase = nphase;
    }
  } else if (warp >= 11) {
    // ------------------------------------------------------------ dQ drain: TMEM -> swizzled staging -> TMA reduce-add
    const int quarter = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + kColDQ;
    uint8_t* stg = sStage + (warp - 11) * 4096;     // [32 q x 32 d] fp32 = 32 rows of 128 bytes, 128B swizzle
    uint4* srow = reinterpret_cast<uint4*>(stg + lane * 128);
    int pp = p0;
    for (int p = 0; p < n_pairs; ++p) {
      ptx::mbar_wait(dq_full, p & 1);
      ptx::tc_fence_after();
      const int q_row = pp * BQ + quarter * 32;
#pragma unroll 1
      for (int hc = 0; hc < 2; ++hc) {               // two 32-column halves of the 64 head-dim columns
        uint32_t v[32];
        ptx::tmem_ld_32x32(lane_addr + hc * 32, v);
        ptx::tmem_ld_wait();
        if (hc == 1) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(dq_free);                 // both halves are in registers / staged
        }
        if (lane == 0) ptx::bulk_wait_read0();       // the previous reduce has read the staging tile
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) srow[j ^ (lane & 7)] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
#ifndef O2_V3_NOREDUCE
          if (q_row < a.N) ptx::tma_reduce_add_3d(&tmap_dq, stg, hc * 32, q_row, bh);
#endif
          ptx::bulk_commit_group();
        }
      }
      if (++pp == n_pairs) pp = 0;
    }
    if (lane == 0) ptx::bulk_wait0();               // every partial has reached the accumulator before the CTA retires
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax: thread = key row r, query columns [64 chalf, +64)
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sc = a.scale_log2;
    const uint64_t sc2 = ptx::pack2(sc, sc);
    const int key = k0 + r;
    const uint32_t drop_key = kDrop ? ptx::attn_drop_key(a.drop, bh) : 0u;
    const uint32_t kbit = kDrop ? (1u << ptx::attn_keep_bit((uint32_t)key)) : 0u;
    const uint32_t sw = (uint32_t)(r & 7);
    int stage = 0;
    uint32_t phase = 0;
    int pp = p0;
    for (int t = 0; t < n_pairs; ++t) {
      const float4* st_l = reinterpret_cast<const float4*>(sStat + stage * 256 + chalf * 64);          // -lse log2e
      const float4* st_d = reinterpret_cast<const float4*>(sStat + stage * 256 + 128 + chalf * 64);    // -delta
      ptx::mbar_wait(&qdo_full[stage], phase);       // the statistics of this query tile are in shared memory
      ptx::mbar_wait(s_full, t & 1);
      ptx::tc_fence_after();
      uint32_t sv_[2][32];
      ptx::tmem_ld_32x32(lane_addr + (uint32_t)(chalf * 64), sv_[0]);
      ptx::tmem_ld_32x32(lane_addr + (uint32_t)(chalf * 64 + 32), sv_[1]);
      uint32_t words[2] = {0u, 0u};
      if (kDrop) {
        words[0] = ptx::attn_keep_word(a.drop, drop_key, (uint32_t)(pp * BQ + chalf * 64 + lane), (uint32_t)(key >> 5));
        words[1] = ptx::attn_keep_word(a.drop, drop_key, (uint32_t)(pp * BQ + chalf * 64 + 32 + lane), (uint32_t)(key >> 5));
      }
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(s_loaded);                    // the tensor pipe may overwrite S^T (with the tile after the next)
      uint32_t pk[2][16];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
#ifdef O2_V3_NOMATH                                   // timing experiment only (wrong results): no softmax arithmetic
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[jj][i] = sv_[jj][i] ^ sv_[jj][i + 16];
        continue;
#endif
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 l4 = st_l[jj * 8 + (i >> 2)];
          const uint64_t X0 = ptx::fma2(ptx::pack2u(sv_[jj][i], sv_[jj][i + 1]), sc2, ptx::pack2(l4.x, l4.y));
          const uint64_t X1 = ptx::fma2(ptx::pack2u(sv_[jj][i + 2], sv_[jj][i + 3]), sc2, ptx::pack2(l4.z, l4.w));
#ifdef O2_V3_NOEXP                                    // timing experiment only (wrong results): no exponentials
          const uint64_t P0 = X0, P1 = X1;
#else
          const uint64_t P0 = (((i >> 1) % kPolyV3) == kPolyV3 - 1) ? ptx::exp2_pair<true>(X0) : ptx::exp2_pair<false>(X0);
          const uint64_t P1 = ((((i >> 1) + 1) % kPolyV3) == kPolyV3 - 1) ? ptx::exp2_pair<true>(X1) : ptx::exp2_pair<false>(X1);
#endif
          pk[jj][i >> 1] = ptx::pack_bf16x2_pair(P0);
          pk[jj][(i >> 1) + 1] = ptx::pack_bf16x2_pair(P1);
        }
      }
      ptx::mbar_wait(dp_full, t & 1);
      ptx::tc_fence_after();
      ptx::mbar_wait(&ds_free[t & 1], ((t >> 1) & 1) ^ 1);     // dQ(t - 2) has read this dS^T set
      uint8_t* ds_base = sdS + ((t & 1) * 2 + chalf) * kT + r * 128;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const uint32_t col = kColDP + (uint32_t)(chalf * 64 + jj * 32);
        uint32_t dv_[32], dk[16];
        ptx::tmem_ld_32x32(lane_addr + col, dv_);
        ptx::tmem_ld_wait();
#ifdef O2_V3_NOMATH
#pragma unroll
        for (int i = 0; i < 16; ++i) dk[i] = dv_[i] ^ dv_[i + 16] ^ pk[jj][i];
#else
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 d4 = st_d[jj * 8 + (i >> 2)];
          uint64_t G0 = ptx::pack2u(dv_[i], dv_[i + 1]), G1 = ptx::pack2u(dv_[i + 2], dv_[i + 3]);
          if (kDrop) {
            float g0, g1, g2, g3;
            ptx::unpack2(G0, g0, g1);
            ptx::unpack2(G1, g2, g3);
            g0 = (__shfl_sync(0xffffffffu, words[jj], i) & kbit) ? g0 * a.drop.inv_keep : 0.f;
            g1 = (__shfl_sync(0xffffffffu, words[jj], i + 1) & kbit) ? g1 * a.drop.inv_keep : 0.f;
            g2 = (__shfl_sync(0xffffffffu, words[jj], i + 2) & kbit) ? g2 * a.drop.inv_keep : 0.f;
            g3 = (__shfl_sync(0xffffffffu, words[jj], i + 3) & kbit) ? g3 * a.drop.inv_keep : 0.f;
            G0 = ptx::pack2(g0, g1);
            G1 = ptx::pack2(g2, g3);
          }
          G0 = ptx::add2(G0, ptx::pack2(d4.x, d4.y));
          G1 = ptx::add2(G1, ptx::pack2(d4.z, d4.w));
          const uint32_t w0 = pk[jj][i >> 1], w1 = pk[jj][(i >> 1) + 1];      // bf16 P pairs -> fp32
          const uint64_t Pa = ptx::pack2u(w0 << 16, w0 & 0xFFFF0000u), Pb = ptx::pack2u(w1 << 16, w1 & 0xFFFF0000u);
          dk[i >> 1] = ptx::pack_bf16x2_pair(ptx::mul2(Pa, G0));
          dk[(i >> 1) + 1] = ptx::pack_bf16x2_pair(ptx::mul2(Pb, G1));
        }
#endif
        ptx::tmem_st_32x16(lane_addr + col, dk);
#pragma unroll
        for (int j = 0; j < 4; ++j)                  // this chunk = 64 bytes of the dS^T row (atom chalf), 128B swizzle
          *reinterpret_cast<uint4*>(ds_base + ((((uint32_t)(jj * 4 + j)) ^ sw) << 4)) =
              make_uint4(dk[4 * j], dk[4 * j + 1], dk[4 * j + 2], dk[4 * j + 3]);
      }
      ptx::fence_proxy_async_smem();
      // P^T (kept probabilities under dropout) -> its own columns, K-major: query q of the tile at column q / 2
      if (t > 0) {
        ptx::mbar_wait(pt_free, (t - 1) & 1);        // dV(t - 1) has read the previous P^T
        ptx::tc_fence_after();
      }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        if (kDrop) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint32_t m0 = (__shfl_sync(0xffffffffu, words[jj], i) & kbit) ? 0x0000FFFFu : 0u;
            const uint32_t m1 = (__shfl_sync(0xffffffffu, words[jj], i + 1) & kbit) ? 0xFFFF0000u : 0u;
            pk[jj][i >> 1] &= (m0 | m1);
          }
        }
        ptx::tmem_st_32x16(lane_addr + kColPT + (uint32_t)(chalf * 32 + jj * 16), pk[jj]);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(pd_full);
      if (++stage == kSt) { stage = 0; phase ^= 1; }
      if (++pp == n_pairs) pp = 0;
    }
    // epilogue: warps 2-5 drain dK (scaled), warps 6-9 dV
    ptx::mbar_wait(dkv_done, 0);
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
