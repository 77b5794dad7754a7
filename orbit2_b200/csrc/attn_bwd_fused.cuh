// One-pass flash-attention backward for head dim 64 (bf16 operands, fp32 accumulation) -- textually included by
// attn_tc.cu inside its anonymous namespace (shares BwdArgs, the descriptor helpers, neg_split3 and the dropout helpers).
//
// reference semantics: components/attention.py:54-78 (autograd backward of softmax(q k^T hd^-0.5) v, attn_drop included).
//
// The two-kernel backward (attn_bwd_dkv_kernel + attn_bwd_dq_kernel) computes S = Q K^T and dP = dO V^T twice (7 GEMMs and
// two softmax passes for 5 algorithmic GEMMs and one).  Here every (128-key tile, 64-query sub-tile) block is visited ONCE:
//     S^T = K Q^T - 1 (lse/scale)^T        dP^T = V dO^T - 1 delta^T            (tensor memory, rank-1 statistics updates)
//     P^T = exp2(S^T c)   dS^T = P^T o dP^T                                     (softmax threads, one key row each)
//     dV += P^T dO        dK += dS^T Q                                          (TS-MMA, accumulators stay in TMEM)
//     dQ[q tile] = dS K                                                         (SS-MMA once per 128 queries: A = dS^T as
//                                                                                written by the softmax threads = dS in
//                                                                                MN-major form, B = K in MN-major form)
// A CTA owns one 128-key tile of one (batch, head) for the whole query range, so dK / dV are exact single-owner sums; its
// dQ partial of every 128-query tile (128 x 64 fp32) is added into the fp32 accumulator dq_accum [B, heads, N, 64] by TMA
// reduce (cp.reduce.async.bulk.tensor .add.f32: the additions run in the L2, 4.8 TB/s measured next to an equal load
// stream, tools/red_bench.cu) and attn_dq_finish_kernel writes scale * dq_accum as bf16.  The summation order over key
// tiles is therefore not fixed: dQ is reproducible to fp32 rounding, not bit for bit (the two-kernel path stays available
// as the deterministic option, O2_ATTN_BWD_TWO_PASS=1).
// CTAs of one (batch, head) start their query loop at different 128-query tiles (tile = key tile index, then wrapping), so
// that the concurrent reduces of a head hit different accumulator tiles.
//
// warp roles (480 threads):  0 TMA producer (K tile once, then the Q / dO ring)    1 tcgen05 issuer (+ TMEM owner)
//                            2-9 softmax (thread = key row, 32 of the 64 query columns of the sub-tile)
//                            10 statistics (-lse/scale, -delta as 3-term bf16 splits -> B operand of the rank-1 updates)
//                            11-14 dQ drain (TMEM -> swizzled staging -> TMA reduce)
// TMEM columns: buffer b: S^T [128 b, +64) dP^T [128 b + 64, +64) | dK [256,320) dV [320,384) | K [384,416) V [416,448)
//               | dQ [448,512).   K and V live in TMEM (packed bf16) so S^T / dP^T are TS-MMAs (32 instead of 58 cycles).
// shared memory (224 KiB): Q / dO ring 3 x 2 x 16 | statistics tile 16 | ones tile 16 | K tile 16 | dS^T 2 sets x 2 x 16
//               | dQ staging 16.
constexpr int kFusedThreads = 480;
constexpr int kFusedStages = 3;
constexpr uint32_t kFusedSmem = (2 * kFusedStages + 3 + 4 + 1) * kTileBytes + 1024 + 256;

template <bool kDrop>
__global__ void __launch_bounds__(kFusedThreads, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                      const __grid_constant__ CUtensorMap tmap_dq, const BwdArgs a) {
  constexpr int kStagesB = kFusedStages;
  constexpr uint32_t kT = kTileBytes;
  constexpr uint32_t kColDK = 256, kColDV = 320, kColK = 384, kColV = 416, kColDQ = 448;
  static_assert(!kDrop || kStagesB <= 3, "the raw delta copy of the dropout variant lives in k-slice 3 of the statistics tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                               // kStagesB tiles of 128 queries
  uint8_t* sdO = sQ + kStagesB * kT;                // kStagesB tiles
  uint8_t* sStat = sdO + kStagesB * kT;             // [128 q x 64]: k-slice `stage` = -(split lse/scale | split delta)
  uint8_t* sOnes = sStat + kT;                      // [128 k x 64]: k-slice 0 ones at positions 0..2, k-slice 1 at 8..10
  uint8_t* sK = sOnes + kT;                         // K tile [128 k x 64 d]: B operand (MN-major) of dQ = dS K
  uint8_t* sdS = sK + kT;                           // 2 sets x 2 tiles [128 k x 64 q] bf16: dS^T rows = dS, MN-major
  uint8_t* sStage = sdS + 4 * kT;                   // 4 drain warps x [32 q x 32 d] fp32, 128B-swizzled
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + kT);
  uint64_t* kv_ready = bars;                        // 1 (256 arrivals: K and V are in TMEM)
  uint64_t* k_full = kv_ready + 1;                  // 1 (TMA: K tile in shared memory)
  uint64_t* qdo_full = k_full + 1;                  // kStagesB (2 arrivals: TMA expect_tx + statistics)
  uint64_t* qdo_empty = qdo_full + kStagesB;
  uint64_t* sd_full = qdo_empty + kStagesB;         // 2 (per buffer)
  uint64_t* pd_full = sd_full + 2;                  // 2 (256 arrivals)
  uint64_t* dkv_done = pd_full + 2;                 // 1
  uint64_t* dq_full = dkv_done + 1;                 // 1: the dQ partial of a 128-query tile is complete in TMEM
  uint64_t* dq_free = dq_full + 1;                  // 1 (128 arrivals): the drain warps hold it in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / a.heads, h = bh % a.heads;
  const int kt = blockIdx.x;
  const int k0 = kt * BKV;
  const int n_pairs = (a.N + BQ - 1) / BQ;          // 128-query tiles; every tile is processed as two 64-query sub-tiles
  const int n_sub = 2 * n_pairs;                    // (rows >= N arrive as zeros and contribute nothing)
#ifdef O2_FUSED_NO_ROT
  const int p0 = 0;
#else
  const int p0 = kt % n_pairs;                      // first 128-query tile of this CTA (see the header comment)
#endif

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::prefetch_tmap(&tmap_do);
    ptx::prefetch_tmap(&tmap_dq);
    ptx::mbar_init(kv_ready, 256);
    ptx::mbar_init(k_full, 1);
    for (int s = 0; s < kStagesB; ++s) {
      ptx::mbar_init(&qdo_full[s], 2);
      ptx::mbar_init(&qdo_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      ptx::mbar_init(&sd_full[t], 1);
      ptx::mbar_init(&pd_full[t], 256);
    }
    ptx::mbar_init(dkv_done, 1);
    ptx::mbar_init(dq_full, 1);
    ptx::mbar_init(dq_free, 128);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + BKV) {         // ones tile (A operand of the rank-1 statistics updates)
    const int r = threadIdx.x - 64;
    const uint4 ones = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
    uint4* row = reinterpret_cast<uint4*>(sOnes + r * 128);
    row[0 ^ (r & 7)] = ones;  row[1 ^ (r & 7)] = zero;       // k-slice 0: positions 0..2
    row[2 ^ (r & 7)] = zero;  row[3 ^ (r & 7)] = ones;       // k-slice 1: positions 8..10
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::mbar_expect_tx(k_full, kT);
      ptx::tma_load_4d(sK, &tmap_qkv, k_full, 0, a.heads + h, k0, b);
      int stage = 0;
      uint32_t phase = 0;
      int pp = p0;
      for (int i = 0; i < n_pairs; ++i) {
        ptx::mbar_wait(&qdo_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&qdo_full[stage], 2 * kT);
        ptx::tma_load_4d(sQ + stage * kT, &tmap_qkv, &qdo_full[stage], 0, h, pp * BQ, b);
        ptx::tma_load_4d(sdO + stage * kT, &tmap_do, &qdo_full[stage], 0, h, pp * BQ, b);
        if (++stage == kStagesB) { stage = 0; phase ^= 1; }
        if (++pp == n_pairs) pp = 0;
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------ statistics warp: the second arrival on qdo_full
    int stage = 0;
    uint32_t phase = 0;
    const float* lse = a.lse + ((size_t)b * a.heads + h) * a.N;
    const float* dlt = a.delta + ((size_t)b * a.heads + h) * a.N;
    const float inv_scale = 1.f / a.scale;
    float xa[4], ya[4], xb[4], yb[4];
    auto tile_of = [&](int i) { int t = p0 + i; return t >= n_pairs ? t - n_pairs : t; };
    auto fetch_stats = [&](int i, float (&xs)[4], float (&ys)[4]) {
      float lv[4], dv[4];
      const int base = (i < n_pairs ? tile_of(i) : 0) * BQ;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = base + e * 32 + lane;
        const bool ok = row < a.N && i < n_pairs;
        lv[e] = ok ? __ldg(lse + row) : 0.f;
        dv[e] = ok ? __ldg(dlt + row) : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = base + e * 32 + lane;
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
        if (kDrop) reinterpret_cast<float*>(dst + (6 ^ (idx & 7)))[stage] = -ys[e];     // raw -delta for the softmax threads
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&qdo_full[stage]);
      if (++stage == kStagesB) { stage = 0; phase ^= 1; }
    };
    fetch_stats(0, xa, ya);
    fetch_stats(1, xb, yb);
    for (int i = 0; i < n_pairs; i += 2) {
      produce(xa, ya);
      fetch_stats(i + 2, xa, ya);
      if (i + 1 < n_pairs) {
        produce(xb, yb);
        fetch_stats(i + 3, xb, yb);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ tcgen05 issuer (warp-uniform loop, one elected lane)
    const uint32_t idesc_s = ptx::umma_idesc_bf16(BKV, BS, 0, 0);   // S^T / dP^T: N = 64 queries
    const uint32_t idesc_g = ptx::umma_idesc_bf16(BKV, kHD, 0, 1);  // dV / dK: A (TMEM), B = dO / Q (MN-major)
    const uint32_t idesc_q = ptx::umma_idesc_bf16(BQ, kHD, 1, 1);   // dQ: A = dS (MN-major: rows = keys), B = K (MN-major)
    const uint64_t dq0 = ptx::umma_smem_desc(ptx::smem_u32(sQ), 16, 1024);
    const uint64_t ddo0 = ptx::umma_smem_desc(ptx::smem_u32(sdO), 16, 1024);
    const uint64_t dqm0 = ptx::umma_smem_desc(ptx::smem_u32(sQ), 8192, 1024);    // Q as MN-major B
    const uint64_t ddom0 = ptx::umma_smem_desc(ptx::smem_u32(sdO), 8192, 1024);  // dO as MN-major B
    const uint64_t dstat0 = ptx::umma_smem_desc(ptx::smem_u32(sStat), 16, 1024);
    const uint64_t dones = ptx::umma_smem_desc(ptx::smem_u32(sOnes), 16, 1024);
    const uint64_t dkm0 = ptx::umma_smem_desc(ptx::smem_u32(sK), 8192, 1024);    // K as MN-major B (64 columns: one atom)
    // dS: M = 128 queries = two 64-query atoms (the two sub-tile tiles, kT apart), K = keys in 8-row groups of 1024 bytes
    const uint64_t dsd0 = ptx::umma_smem_desc(ptx::smem_u32(sdS), kT, 1024);
    auto issue_sd = [&](int buf, int stage, int half) {
      if (ptx::elect_one()) {
        const uint64_t qa = desc_add(dq0, stage * kT + half * kHalfBytes);
        const uint64_t da = desc_add(ddo0, stage * kT + half * kHalfBytes);
        const uint64_t st = desc_add(dstat0, half * kHalfBytes + stage * 32);
        const uint32_t ds_ = tmem_base + buf * 128;
        ptx::umma_ts(ds_, tmem_base + kColK, qa, idesc_s, 0u);
#pragma unroll
        for (int k = 1; k < 4; ++k) ptx::umma_ts_acc(ds_, tmem_base + kColK + k * 8, desc_add(qa, k * 32), idesc_s);
        ptx::umma_ss_acc(ds_, dones, st, idesc_s);                                         // - lse / scale
        ptx::umma_ts(ds_ + 64, tmem_base + kColV, da, idesc_s, 0u);
#pragma unroll
        for (int k = 1; k < 4; ++k) ptx::umma_ts_acc(ds_ + 64, tmem_base + kColV + k * 8, desc_add(da, k * 32), idesc_s);
        if (!kDrop) ptx::umma_ss_acc(ds_ + 64, desc_add(dones, 32), st, idesc_s);         // - delta
        ptx::umma_commit(&sd_full[buf]);
      }
      __syncwarp();
    };
    ptx::mbar_wait(kv_ready, 0);
    ptx::mbar_wait(k_full, 0);
    ptx::mbar_wait(&qdo_full[0], 0);
    ptx::tc_fence_after();
    issue_sd(0, 0, 0);
    issue_sd(1, 0, 1);
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
        if (u + 1 == n_sub) ptx::umma_commit(dkv_done);
      }
      __syncwarp();
#ifndef O2_FUSED_DQ_EARLY
      if (u + 2 < n_sub) {                            // this buffer's next sub-tile (u + 2) lives in the next 128-query tile
        if (buf == 0) {                               // first touch of that tile
          ptx::mbar_wait(&qdo_full[nstage], nphase);
          ptx::tc_fence_after();
        }
        issue_sd(buf, nstage, buf);
      }
#endif
      if (buf == 1) {                                 // both halves of this 128-query tile are done
        if (ptx::elect_one()) ptx::umma_commit(&qdo_empty[stage]);     // Q / dO stage: every MMA that reads it is issued
        __syncwarp();
        const int p = u >> 1;
#ifndef O2_FUSED_NO_DQ
        if (p > 0) {                                  // the drain warps hold the previous partial in registers
          ptx::mbar_wait(dq_free, (p - 1) & 1);
          ptx::tc_fence_after();
        }
#endif
        if (ptx::elect_one()) {
#ifndef O2_FUSED_NO_DQ
          // dQ[128 q x 64 d] = dS[128 q x 128 k] K[128 k x 64 d].  It is issued AFTER S^T / dP^T of sub-tile u + 2, so it
          // never delays the next softmax; sd_full(u + 4) -- the hand-off that lets the softmax threads overwrite this
          // dS set -- is committed after these MMAs and therefore covers them.
          const uint64_t a0 = desc_add(dsd0, (p & 1) * 2 * kT);
          ptx::umma_ss_first(tmem_base + kColDQ, a0, dkm0, idesc_q);
#pragma unroll
          for (int k = 1; k < BKV / 16; ++k)
            ptx::umma_ss_acc(tmem_base + kColDQ, desc_add(a0, k * 2048), desc_add(dkm0, k * 2048), idesc_q);
          ptx::umma_commit(dq_full);
#endif
        }
        __syncwarp();
      }
#ifdef O2_FUSED_DQ_EARLY
      if (u + 2 < n_sub) {
        if (buf == 0) {
          ptx::mbar_wait(&qdo_full[nstage], nphase);
          ptx::tc_fence_after();
        }
        issue_sd(buf, nstage, buf);
      }
#endif
      if (buf == 1) {
        stage = nstage;
        phase = nphase;
      }
    }
  } else if (warp >= 11) {
    // ------------------------------------------------------------ dQ drain: TMEM -> swizzled staging -> TMA reduce-add
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may touch = query rows [32 quarter, +32)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + kColDQ;
    uint8_t* stg = sStage + (warp - 11) * 4096;     // [32 q x 32 d] fp32 = 32 rows of 128 bytes, 128B swizzle
    uint4* srow = reinterpret_cast<uint4*>(stg + lane * 128);
    int pp = p0;
#ifdef O2_FUSED_NO_DQ
    for (int p = 0; p < 0; ++p) {
#else
    for (int p = 0; p < n_pairs; ++p) {
#endif
      ptx::mbar_wait(dq_full, p & 1);
      ptx::tc_fence_after();
      uint32_t v0[32], v1[32];
      ptx::tmem_ld_32x32(lane_addr, v0);
      ptx::tmem_ld_32x32(lane_addr + 32, v1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(dq_free);
      const int q_row = pp * BQ + quarter * 32;
      if (lane == 0) ptx::bulk_wait_read0();        // the previous reduce has read the staging tile
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) srow[j ^ (lane & 7)] = make_uint4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
#ifndef O2_FUSED_NO_REDUCE
        if (q_row < a.N) ptx::tma_reduce_add_3d(&tmap_dq, stg, 0, q_row, bh);
#endif
        ptx::bulk_commit_group();
        ptx::bulk_wait_read0();
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) srow[j ^ (lane & 7)] = make_uint4(v1[4 * j], v1[4 * j + 1], v1[4 * j + 2], v1[4 * j + 3]);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
#ifndef O2_FUSED_NO_REDUCE
        if (q_row < a.N) ptx::tma_reduce_add_3d(&tmap_dq, stg, 32, q_row, bh);
#endif
        ptx::bulk_commit_group();
      }
      if (++pp == n_pairs) pp = 0;
    }
    if (lane == 0) ptx::bulk_wait0();               // every partial has reached the accumulator before the CTA retires
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax: thread = one key row, 32 of the 64 query columns
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;            // query columns [32 chalf, +32) of the sub-tile
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sc = a.scale_log2;
    const uint64_t sc2 = ptx::pack2(sc, sc);
    const int key = k0 + r;
    {   // K row (warps 2-5) / V row (warps 6-9) of this thread's key -> its TMEM lane
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
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(kv_ready);
    }
    const uint32_t drop_key = kDrop ? ptx::attn_drop_key(a.drop, bh) : 0u;
    // this thread's 64 bytes (32 queries) of a dS^T row: chunks 4 chalf .. 4 chalf + 3 of the 128-byte row, 128B swizzle
    const uint32_t ds_row = r * 128;
    const uint32_t sw = (uint32_t)(r & 7);
    int dstage = 0;                              // stage of the 128-query tile that holds sub-tile u
    int pp = p0;                                 // its tile index
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
        // Keep decisions of this thread's key for its 32 queries as ONE word in the bit order the PRMT mask expansion wants
        // (pi = attn_keep_bit): lane l builds the keep word of query qb + pi^-1(l) (32 keys of this warp's key block), the
        // 32 x 32 bit matrix is transposed across the warp with five butterfly shuffles, and the row of key `key` (bit
        // position pi(key)) is fetched from the lane that holds it: 6 shuffles per 32 elements (the first version
        // shuffled every word to every lane: 32), masks applied as AND-masks on packed values, math back in fp32x2.
        const int qb = pp * BQ + buf * BS + chalf * 32;
        const uint32_t my_word = ptx::attn_keep_word(a.drop, drop_key, (uint32_t)(qb + (int)ptx::attn_keep_bit_inv((uint32_t)lane)),
                                                     (uint32_t)(key >> 5));
        uint32_t T = ptx::warp_transpose32(my_word, lane);
        T = __shfl_sync(0xffffffffu, T, (int)ptx::attn_keep_bit((uint32_t)key));
        drop_apply_f32(dv_, T);                                     // dP^T through the mask (dropped -> +0)
        const float* dl = reinterpret_cast<const float*>(sStat) + dstage;      // -delta of the tile's 128 queries
        const uint64_t ik2 = ptx::pack2(a.drop.inv_keep, a.drop.inv_keep);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t X = ptx::mul2(ptx::pack2u(sv_[i], sv_[i + 1]), sc2);
          const uint64_t Pq = ptx::exp2_pair<false>(X);
          const int qi = buf * BS + chalf * 32 + i;                // row of the 128-query statistics tile
          const uint64_t nd = ptx::pack2(dl[(qi * 128 + ((6 ^ (qi & 7)) * 16)) >> 2],
                                         dl[((qi + 1) * 128 + ((6 ^ ((qi + 1) & 7)) * 16)) >> 2]);
          const uint64_t G = ptx::fma2(ptx::pack2u(dv_[i], dv_[i + 1]), ik2, nd);    // M dP / keep - delta
          pk[i >> 1] = ptx::pack_bf16x2_pair(Pq);
          dk[i >> 1] = ptx::pack_bf16x2_pair(ptx::mul2(Pq, G));
        }
        drop_apply_bf16x2(pk, T);                                   // dV uses the masked probabilities
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t X = ptx::mul2(ptx::pack2u(sv_[i], sv_[i + 1]), sc2);
          const uint64_t Pq = (((i >> 1) % kPolyFused) == kPolyFused - 1) ? ptx::exp2_pair<true>(X) : ptx::exp2_pair<false>(X);
          pk[i >> 1] = ptx::pack_bf16x2_pair(Pq);
          dk[i >> 1] = ptx::pack_bf16x2_pair(ptx::mul2(Pq, ptx::pack2u(dv_[i], dv_[i + 1])));
        }
      }
      ptx::tmem_st_32x16(st_addr, pk);
      ptx::tmem_st_32x16(dp_addr, dk);
      {   // dS^T row piece -> shared memory (A operand of dQ = dS K); set = tile parity, tile = sub-tile half
        uint8_t* base = sdS + (((u >> 1) & 1) * 2 + buf) * kT + ds_row;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(base + ((((uint32_t)(chalf * 4 + j)) ^ sw) << 4)) =
              make_uint4(dk[4 * j], dk[4 * j + 1], dk[4 * j + 2], dk[4 * j + 3]);
      }
      ptx::fence_proxy_async_smem();
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&pd_full[buf]);
      if (buf == 1) {
        if (++dstage == kStagesB) dstage = 0;
        if (++pp == n_pairs) pp = 0;
      }
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

// dQ = scale * dq_accum (fp32 [B, heads, N, 64]) -> bf16 into the q slot of dqkv [B, N, 3, heads, 64]; 8 threads per row
__global__ void attn_dq_finish_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, int B, int N, int heads,
                                      float scale) {
  const long long gid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3;     // (b, h, n) row
  const int sub = threadIdx.x & 7;
  const long long total = (long long)B * heads * N;
  if (gid >= total) return;
  const int n = (int)(gid % N);
  const long long bh_ = gid / N;
  const int h = (int)(bh_ % heads), b = (int)(bh_ / heads);
  const float4 x0 = *reinterpret_cast<const float4*>(acc + gid * 64 + sub * 8);
  const float4 x1 = *reinterpret_cast<const float4*>(acc + gid * 64 + sub * 8 + 4);
  uint4 w;
  w.x = pack_bf16x2(x0.x * scale, x0.y * scale);
  w.y = pack_bf16x2(x0.z * scale, x0.w * scale);
  w.z = pack_bf16x2(x1.x * scale, x1.y * scale);
  w.w = pack_bf16x2(x1.z * scale, x1.w * scale);
  *reinterpret_cast<uint4*>(dqkv + ((((size_t)b * N + n) * 3 + 0) * heads + h) * 64 + sub * 8) = w;
}
