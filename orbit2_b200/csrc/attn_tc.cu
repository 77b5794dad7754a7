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
constexpr int kStagesF = 3;
constexpr int kThreadsF = 320;
constexpr uint32_t kTileBytes = BQ * kHD * 2;   // 16 KiB (Q, K and V tiles alike)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;

struct FwdArgs {
  __nv_bfloat16* out;
  float* lse;
  int B, N, heads, n_kv;
  float scale_log2;   // softmax scale * log2(e)
};

struct FwdSmem {
  // barriers live after the tiles; layout computed by hand below
};

constexpr uint32_t kFwdSmemBytes = (2 + 2 * kStagesF) * kTileBytes + 1024 + 256;

__global__ void __launch_bounds__(kThreadsF, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // 2 tiles
  uint8_t* sK = sQ + 2 * kTileBytes;                // kStagesF tiles
  uint8_t* sV = sK + kStagesF * kTileBytes;         // kStagesF tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStagesF * kTileBytes);
  uint64_t* q_full = bars;                          // 1
  uint64_t* k_full = q_full + 1;                    // kStagesF
  uint64_t* k_empty = k_full + kStagesF;
  uint64_t* v_full = k_empty + kStagesF;
  uint64_t* v_empty = v_full + kStagesF;
  uint64_t* s_full = v_empty + kStagesF;            // 2
  uint64_t* p_full = s_full + 2;                    // 2
  uint64_t* o_done = p_full + 2;                    // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / a.heads, h = bh % a.heads;
  const int q0 = blockIdx.x * 2 * BQ;
  const int n_kv = a.n_kv;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < kStagesF; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], 1);
      ptx::mbar_init(&v_full[s], 1);
      ptx::mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      ptx::mbar_init(&s_full[t], 1);
      ptx::mbar_init(&p_full[t], 128);
      ptx::mbar_init(&o_done[t], 1);
    }
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
      ptx::mbar_expect_tx(q_full, 2 * kTileBytes);
      ptx::tma_load_4d(sQ, &tmap_qkv, q_full, 0, h, q0, b);
      ptx::tma_load_4d(sQ + kTileBytes, &tmap_qkv, q_full, 0, h, q0 + BQ, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        ptx::mbar_wait(&k_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&k_full[stage], kTileBytes);
        ptx::tma_load_4d(sK + stage * kTileBytes, &tmap_qkv, &k_full[stage], 0, a.heads + h, j * BKV, b);
        ptx::mbar_wait(&v_empty[stage], phase ^ 1);
        ptx::mbar_expect_tx(&v_full[stage], kTileBytes);
        ptx::tma_load_4d(sV + stage * kTileBytes, &tmap_qkv, &v_full[stage], 0, 2 * a.heads + h, j * BKV, b);
        if (++stage == kStagesF) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ tcgen05 issuer
    if (lane == 0) {
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t idesc_o = ptx::umma_idesc_bf16(BQ, kHD, 0, 1);   // A = P (TMEM, K-major), B = V (MN-major)
      const uint32_t sq = ptx::smem_u32(sQ), sk = ptx::smem_u32(sK), sv = ptx::smem_u32(sV);
      auto issue_s = [&](int t, int kstage) {
        const uint32_t qa = sq + t * kTileBytes, ka = sk + kstage * kTileBytes;
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k)
          ptx::umma_ss(tmem_base + t * 128, ptx::umma_smem_desc(qa + k * 32, 16, 1024),
                       ptx::umma_smem_desc(ka + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit(&s_full[t]);
      };
      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&k_full[0], 0);
      ptx::tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      ptx::umma_commit(&k_empty[0]);
      int kstage = 1 % kStagesF;
      uint32_t kphase = (kStagesF == 1) ? 1 : 0;
      int vstage = 0;
      uint32_t vphase = 0;
      for (int j = 0; j < n_kv; ++j) {
        ptx::mbar_wait(&v_full[vstage], vphase);
        const bool more = (j + 1 < n_kv);
        for (int t = 0; t < 2; ++t) {
          ptx::mbar_wait(&p_full[t], j & 1);
          ptx::tc_fence_after();
          const uint32_t va = sv + vstage * kTileBytes;
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)
            ptx::umma_ts(tmem_base + 256 + t * 64, tmem_base + t * 128 + k * 8,
                         ptx::umma_smem_desc(va + k * 2048, 8192, 1024), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          ptx::umma_commit(&o_done[t]);
          if (more) {
            if (t == 0) {
              ptx::mbar_wait(&k_full[kstage], kphase);
              ptx::tc_fence_after();
            }
            issue_s(t, kstage);
            if (t == 1) {
              ptx::umma_commit(&k_empty[kstage]);
              if (++kstage == kStagesF) { kstage = 0; kphase ^= 1; }
            }
          }
        }
        ptx::umma_commit(&v_empty[vstage]);
        if (++vstage == kStagesF) { vstage = 0; vphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / correction / epilogue
    const int t = (warp - 2) >> 2;          // query tile of this warp group
    const int quarter = warp & 3;           // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;      // row inside the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t s_addr = lane_addr + t * 128;
    const uint32_t o_addr = lane_addr + 256 + t * 64;
    const float sc = a.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    const int kv_tail = a.N - (n_kv - 1) * BKV;   // valid keys in the last tile (1..128)
    for (int j = 0; j < n_kv; ++j) {
      ptx::mbar_wait(&s_full[t], j & 1);
      ptx::tc_fence_after();
      const int valid = (j == n_kv - 1) ? kv_tail : BKV;
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BKV / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(s_addr + c * 32, v);
        ptx::tmem_ld_wait();
        if (valid == BKV) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_ref, mx * sc);
      const bool need = __any_sync(0xffffffffu, m_new - m_ref > kRescaleThreshold);
      if (need) {
        const float f = ptx::ex2(m_ref - m_new);     // 0 on the first tile (m_ref = -inf)
        l *= f;
        m_ref = m_new;
        if (j > 0) {
          ptx::mbar_wait(&o_done[t], (j - 1) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < kHD / 32; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(o_addr + c * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            ptx::tmem_st_32x32(o_addr + c * 32, v);
          }
        }
      }
      // pass 2: p = exp2(s * sc - m_ref), row sum, bf16 P written over the head of S
      const float neg_m = -m_ref;
#pragma unroll 1
      for (int c = 0; c < BKV / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(s_addr + c * 32, v);
        ptx::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ptx::ex2(fmaf(__uint_as_float(v[i]), sc, neg_m));
          float p1 = ptx::ex2(fmaf(__uint_as_float(v[i + 1]), sc, neg_m));
          if (valid != BKV) {
            if (c * 32 + i >= valid) p0 = 0.f;
            if (c * 32 + i + 1 >= valid) p1 = 0.f;
          }
          l += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        ptx::tmem_st_32x16(s_addr + c * 16, pk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[t]);
    }
    // epilogue: O / l -> bf16, lse
    ptx::mbar_wait(&o_done[t], (n_kv - 1) & 1);
    ptx::tc_fence_after();
    const int row = q0 + t * BQ + r;
    const float inv = 1.f / l;
    uint32_t o[2][32];
    ptx::tmem_ld_32x32(o_addr, o[0]);
    ptx::tmem_ld_32x32(o_addr + 32, o[1]);
    ptx::tmem_ld_wait();
    if (row < a.N) {
      __nv_bfloat16* op = a.out + (((size_t)b * a.N + row) * a.heads + h) * kHD;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[c][i]) * inv, __uint_as_float(o[c][i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[c][i + 2]) * inv, __uint_as_float(o[c][i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[c][i + 4]) * inv, __uint_as_float(o[c][i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[c][i + 6]) * inv, __uint_as_float(o[c][i + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c * 32 + i) = u;
        }
      a.lse[((size_t)b * a.heads + h) * a.N + row] = (m_ref + log2f(l)) * kLn2;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

int make_qkv_tmap(CUtensorMap* tm, const void* qkv, int B, int N, int heads, int hd, int box_rows) {
  uint64_t dims[4] = {(uint64_t)hd, (uint64_t)(3 * heads), (uint64_t)N, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)hd * 2, (uint64_t)3 * heads * hd * 2, (uint64_t)N * 3 * heads * hd * 2};
  uint32_t box[4] = {(uint32_t)hd, 1, (uint32_t)box_rows, 1};
  return o2_make_tmap(tm, qkv, 2, 4, dims, str, box, 1);
}

}  // namespace

int o2_attn_fwd_tc(const void* qkv, void* out, float* lse, int B, int N, int heads, int hd, float scale, cudaStream_t st) {
  O2_REQUIRE(hd == kHD, "attn_fwd_tc: head dim %d not supported (64 only)", hd);
  O2_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "attn_fwd_tc: pointers must be 16-byte aligned");
  O2_REQUIRE((long long)B * heads <= 65535, "attn_fwd_tc: B*heads too large");
  CUtensorMap tm;
  int rc = make_qkv_tmap(&tm, qkv, B, N, heads, hd, BQ);
  if (rc) return rc;
  FwdArgs a;
  a.out = (__nv_bfloat16*)out; a.lse = lse; a.B = B; a.N = N; a.heads = heads;
  a.n_kv = (N + BKV - 1) / BKV;
  a.scale_log2 = scale * kLog2e;
  static bool attr_done = false;
  if (!attr_done) {
    O2_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmemBytes));
    attr_done = true;
  }
  dim3 grid((N + 2 * BQ - 1) / (2 * BQ), B * heads);
  attn_fwd_tc_kernel<<<grid, kThreadsF, kFwdSmemBytes, st>>>(tm, a);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

int o2_attn_bwd_tc(const void*, const void*, const void*, const float*, void*, float*, int, int, int, int, float,
                   cudaStream_t) {
  O2_FAIL(O2_ERR_UNSUPPORTED, "attn_bwd_tc: not built yet");
}
