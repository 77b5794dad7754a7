// tcgen05 / TMEM / TMA GEMM for sm_100a: C[M,N] = op(A)[M,K] op(B)[K,N] (+ fused epilogue), bf16 in,
// fp32 accumulate in tensor memory.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner),
// warps 2..5 = epilogue (one TMEM lane quarter each).  128 x BN output tile, BK = 64 (one
// 128-byte swizzle atom of bf16), STAGES-deep smem ring, two TMEM accumulator stages so the
// epilogue of tile i overlaps the main loop of tile i+1.
//
// Both operands may be K-major (row = M/N index, K contiguous) or MN-major (row = K index, M/N
// contiguous); the latter is what weight-gradient (dY^T X) and data-gradient (dY W) GEMMs need, so
// no transposed copies of activations or weights are ever materialised.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;    // epilogue warps: kEpiWarps / 4 per TMEM lane quarter, each draining a column slice
constexpr int kThreads = 64 + 32 * kEpiWarps;   // + TMA warp + MMA warp
constexpr uint32_t kStageA = BM * BK * 2;  // 16 KiB

struct GemmArgs {
  int M, N, K;
  int dbg;   // -DO2_GEMM_ABLATE builds only (timing experiments, results invalid): O2_GEMM_DBG=1 skips the pre-activation store, 2 the GELU math
  int epi, c_f32, wide_st, wide_ld, tma_st;   // tma_st: bf16 outputs leave through shared memory + TMA stores   // C (and aux_out) / aux rows are 32-byte aligned -> 256-bit stores / loads
  long long ldc, ld_aux, aux_rows, ld_aux_out;
  const float* bias;
  // token-stream dropout / stochastic depth fused into the epilogue (mask function of csrc/dropout.cu, e = row * N + col):
  // BIAS_RES: C = (acc + bias) * m(e) * sample_scale[row / rows_per_sample] + aux;  BIAS_GELU: C = gelu(pre) * m(e);
  // DGELU: C = acc * m(e) * gelu'(aux);  m = keep / (1 - p)
  int drop;                  // 0: none, 1: mask before the residual add, 2: after it (pos_drop)
  uint32_t drop_key, drop_thr16;
  const uint64_t* step_word;
  float drop_inv_keep;
  const float* sample_scale;
  long long rows_per_sample;
  float* colsum;             // O2_EPI_ACCUM with MN-major A: colsum[m] += sum_k A^T[m, k] (bias gradient riding on the wgrad)
  const __nv_bfloat16* aux;
  __nv_bfloat16* aux_out;
  void* C;
  int a_mn, b_mn;
  int num_m_blk, num_n_blk, split_k, kb_total, kb_per_split;
  uint32_t mn_lbo, mn_sbo;  // descriptor strides for MN-major operands (bytes)
};

// PAIR: the CTA pair runs ONE tcgen05.mma.cta_group::2 per k slice (256 x BN tile): a CTA holds its 128 rows of A and only
// HALF of the B tile (32 instead of 48 KiB per stage at BN = 256 -> 6 stages instead of 4, and 64 instead of 96 bytes per
// clock of shared-memory operand reads per SM).
template <int BN, bool PAIR> struct Cfg {
  static constexpr uint32_t kStageB = (PAIR ? BN / 2 : BN) * BK * 2;
  static constexpr uint32_t kStageBytes = kStageA + kStageB;
  static constexpr int kStages = PAIR ? (BN == 256 ? 6 : 8) : (BN == 256 ? 4 : 6);
  static constexpr int kTmemCols = 2 * BN;  // 512 or 256
  static constexpr uint32_t kStagingBytes = kEpiWarps * 2048;   // per epilogue warp: 32 rows x 32 bf16 columns for the TMA store
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/ +
                                         8 * 128 * 4 /*bias*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// One 32-column chunk of one output row.  `sbias` = this warp's bias slice staged in shared memory (broadcast reads), `ax` =
// the chunk's aux values (residual / pre-activation), fetched by the caller BEFORE it waited for the TMEM load so that
// their L2 latency overlaps it (the first version loaded bias and aux here, per 8 columns: the epilogue warps then sat
// in long-scoreboard stalls for 2/3 of the time and the tensor pipe idled at 58 %).
// 32-byte (one full sector) store: with 16-byte stores every lane of a row-per-lane epilogue half-fills a sector per
// request, and the LSU store path -- not the tensor pipe -- bounded the GELU / residual epilogues (ncu: LDS / FADD of
// the next columns waiting on the previous STG's operand read, 32 sectors per request at 50 % efficiency).
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}
// bf16 row piece of 16 columns starting at n (n % 8 == 0): one 256-bit store when it is in range and 32-byte aligned
__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* dst, const float* v, int n, int N, bool wide) {
  const uint4 lo = pack8(v), hi = pack8(v + 8);
  if (wide && n + 16 <= N) {
    st_global_256(dst, lo, hi);
  } else {
    *reinterpret_cast<uint4*>(dst) = lo;
    if (n + 8 < N) *reinterpret_cast<uint4*>(dst + 8) = hi;
  }
}

// erf-GELU and its derivative for a PAIR of values in packed fp32x2 arithmetic (FFMA2 / FMUL2: two lanes per issue slot).
// Same Abramowitz & Stegun 7.1.26 erf as gelu_fast / dgelu_fast of common.cuh (|error| <= 1.5e-7), rearranged so that no
// sign handling or 1 + erf term is left:   t = 1 / (1 + 0.2316419 |x|),  q = 0.5 P(t) e^(-x^2 / 2)  (the 0.5 folded into the
// polynomial coefficients),   gelu(x) = max(x, 0) - |x| q,   Phi(x) = 0.5 + copysign(0.5 - q, x),
// gelu'(x) = Phi(x) + x e^(-x^2 / 2) / sqrt(2 pi).  Per pair: 6 FFMA2 + 4 FMUL2 + 2 LOP3 + 2 FMNMX + 4 MUFU for gelu (the
// scalar form: 30 FMA-pipe instructions + 4 MUFU; ncu of the fc1 shape: FMUL + FFMA = 41 % of all issued instructions,
// epilogue warps at 27 instructions per element and the tensor pipe waiting for them at 57 %).
#define O2_C2(x) ptx::pack2((x), (x))
struct GeluPair {
  uint64_t X, NAX, Q, E;   // x, -|x|, q, e^(-x^2 / 2)
};
__device__ __forceinline__ GeluPair gelu_pair(float x0, float x1) {
  GeluPair r;
  r.X = ptx::pack2(x0, x1);
  r.NAX = ptx::pack2u(__float_as_uint(x0) | 0x80000000u, __float_as_uint(x1) | 0x80000000u);
  const uint64_t D = ptx::fma2(r.NAX, O2_C2(-0.23164190f), O2_C2(1.0f));
  float d0, d1, t0, t1;
  ptx::unpack2(D, d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t T = ptx::pack2(t0, t1);
  uint64_t P = ptx::fma2(T, O2_C2(0.5f * 1.061405429f), O2_C2(0.5f * -1.453152027f));
  P = ptx::fma2(P, T, O2_C2(0.5f * 1.421413741f));
  P = ptx::fma2(P, T, O2_C2(0.5f * -0.284496736f));
  P = ptx::fma2(P, T, O2_C2(0.5f * 0.254829592f));
  P = ptx::mul2(P, T);
  const uint64_t U = ptx::mul2(ptx::mul2(r.X, r.X), O2_C2(-0.72134752044448170f));     // -x^2 / 2 * log2(e)
  float u0, u1;
  ptx::unpack2(U, u0, u1);
  r.E = ptx::pack2(ptx::ex2(u0), ptx::ex2(u1));
  r.Q = ptx::mul2(P, r.E);
  return r;
}
// the same for kP pairs with every step applied to all pairs before the next one: the dependent chain of one pair (fma2 ->
// rcp -> 4 Horner steps -> ...) is ~12 instructions of 4-30 cycles latency each, and with two epilogue warps per scheduler
// the pairs have to overlap INSIDE a warp (ncu: "wait" + scoreboard stalls 58 % of the samples, 0.2 IPC per warp)
template <int kP>
__device__ __forceinline__ void gelu_pairs(float* v) {          // v[2 kP] in place
  uint64_t X[kP], NAX[kP], T[kP], P[kP], E[kP];
#pragma unroll
  for (int k = 0; k < kP; ++k) {
    X[k] = ptx::pack2(v[2 * k], v[2 * k + 1]);
    NAX[k] = ptx::pack2u(__float_as_uint(v[2 * k]) | 0x80000000u, __float_as_uint(v[2 * k + 1]) | 0x80000000u);
  }
#pragma unroll
  for (int k = 0; k < kP; ++k) T[k] = ptx::fma2(NAX[k], O2_C2(-0.23164190f), O2_C2(1.0f));
#pragma unroll
  for (int k = 0; k < kP; ++k) {
    float d0, d1, t0, t1;
    ptx::unpack2(T[k], d0, d1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
    T[k] = ptx::pack2(t0, t1);
  }
#pragma unroll
  for (int k = 0; k < kP; ++k) E[k] = ptx::mul2(ptx::mul2(X[k], X[k]), O2_C2(-0.72134752044448170f));
#pragma unroll
  for (int k = 0; k < kP; ++k) {
    float u0, u1;
    ptx::unpack2(E[k], u0, u1);
    E[k] = ptx::pack2(ptx::ex2(u0), ptx::ex2(u1));
  }
#pragma unroll
  for (int k = 0; k < kP; ++k) P[k] = ptx::fma2(T[k], O2_C2(0.5f * 1.061405429f), O2_C2(0.5f * -1.453152027f));
#pragma unroll
  for (int k = 0; k < kP; ++k) P[k] = ptx::fma2(P[k], T[k], O2_C2(0.5f * 1.421413741f));
#pragma unroll
  for (int k = 0; k < kP; ++k) P[k] = ptx::fma2(P[k], T[k], O2_C2(0.5f * -0.284496736f));
#pragma unroll
  for (int k = 0; k < kP; ++k) P[k] = ptx::fma2(P[k], T[k], O2_C2(0.5f * 0.254829592f));
#pragma unroll
  for (int k = 0; k < kP; ++k) P[k] = ptx::mul2(ptx::mul2(P[k], T[k]), E[k]);      // q
#pragma unroll
  for (int k = 0; k < kP; ++k)
    ptx::unpack2(ptx::fma2(NAX[k], P[k], ptx::pack2(fmaxf(v[2 * k], 0.f), fmaxf(v[2 * k + 1], 0.f))), v[2 * k], v[2 * k + 1]);
}
// (v0, v1) *= gelu'(a0, a1)
__device__ __forceinline__ void dgelu_mul2(float& v0, float& v1, float a0, float a1) {
  const GeluPair p = gelu_pair(a0, a1);
  float h0, h1;
  ptx::unpack2(ptx::fma2(p.Q, O2_C2(-1.0f), O2_C2(0.5f)), h0, h1);              // 0.5 - q >= 0
  const uint64_t S = ptx::pack2u(__float_as_uint(h0) | (__float_as_uint(a0) & 0x80000000u),
                                 __float_as_uint(h1) | (__float_as_uint(a1) & 0x80000000u));
  const uint64_t Dg = ptx::fma2(ptx::mul2(p.X, O2_C2(0.39894228040143268f)), p.E, ptx::add2(S, O2_C2(0.5f)));
  ptx::unpack2(ptx::mul2(ptx::pack2(v0, v1), Dg), v0, v1);
}

// keep / (1 - p) factors of the 16 elements (row, n .. n + 15) folded into v: one 32-bit hash per element pair
__device__ __forceinline__ void drop_apply16(const GemmArgs& g, float (&v)[16], long long row, int n, float s) {
  const uint32_t key = g.drop_key ^ ptx::step_word_mix(g.step_word);
  const unsigned long long pair0 = (unsigned long long)(row * (long long)g.N + n) >> 1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const unsigned long long pair = pair0 + j;
    const uint32_t h = ptx::lowbias32((uint32_t)pair ^ key ^ ((uint32_t)(pair >> 32) * 0x9E3779B1u));
    v[2 * j] = ((h & 0xFFFFu) >= g.drop_thr16) ? v[2 * j] * s : 0.f;
    v[2 * j + 1] = ((h >> 16) >= g.drop_thr16) ? v[2 * j + 1] * s : 0.f;
  }
}

// bf16 outputs through shared memory and TMA.  A row-per-lane `st.global.v8` touches 32 different 128-byte lines per warp
// instruction and costs the L1 data pipe ~64 wavefronts (ncu, fc1 + GELU shape: l1tex LSU wavefronts 65 % busy, two output
// streams = 8 k wavefront cycles per tile against 8 k cycles of MMA -- the epilogue-"bound" shapes were store-path bound;
// packed GELU math and prefetched TMEM loads changed nothing).  Each epilogue warp stages its 32 rows x 32 columns (64 B per
// row, SWIZZLE_64B: 16-byte chunk c of row r at c ^ ((r >> 1) & 3), conflict-free for one row per lane) and one lane hands
// the 2 KiB block to the TMA, which also clips rows >= M / columns >= N.
struct StoreCtx {
  uint8_t* stage;              // this warp's 2 KiB staging block (1 KiB aligned), or nullptr: direct stores
  const CUtensorMap* tm_c;
  const CUtensorMap* tm_aux;
  int row0;                    // global row of lane 0
  int lane;
};
__device__ __forceinline__ void stage_row16p(const StoreCtx& sc, int j0, const uint4& lo, const uint4& hi);
__device__ __forceinline__ void stage_row16(const StoreCtx& sc, int j0, const float* v) {   // columns j0 .. j0 + 15 of the chunk
  stage_row16p(sc, j0, pack8(v), pack8(v + 8));
}
__device__ __forceinline__ void stage_row16p(const StoreCtx& sc, int j0, const uint4& lo, const uint4& hi) {
  const uint32_t swz = (uint32_t)(sc.lane >> 1) & 3u;
  uint8_t* rowp = sc.stage + sc.lane * 64;
  *reinterpret_cast<uint4*>(rowp + ((((uint32_t)j0 >> 3)) ^ swz) * 16) = lo;
  *reinterpret_cast<uint4*>(rowp + ((((uint32_t)j0 >> 3) + 1) ^ swz) * 16) = hi;
}
__device__ __forceinline__ void stage_wait(const StoreCtx& sc) {     // the previous TMA store has read the block
  if (sc.lane == 0) ptx::bulk_wait_read0();
  __syncwarp();
}
__device__ __forceinline__ void stage_store(const StoreCtx& sc, const CUtensorMap* tm, int n0) {
  ptx::fence_proxy_async_smem();
  __syncwarp();
  if (sc.lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(tm), "r"(ptx::smem_u32(sc.stage)), "r"(n0), "r"(sc.row0) : "memory");
    ptx::bulk_commit_group();
  }
}

template <int BN, int EPI, bool C_F32, bool DROP = false>
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& g, const uint32_t (&r)[32], long long row, int n0,
                                               const float* sbias, const uint4 (&ax)[4], const StoreCtx& sc) {
  constexpr bool kBf16Out = (EPI != O2_EPI_ACCUM) && !C_F32;
  const bool tma = kBf16Out && sc.stage != nullptr;
  uint4 keep[4];                                       // packed GELU outputs wait here while the pre-activation block is stored
  if (tma) {
    if (n0 >= g.N) return;                             // warp-uniform
    stage_wait(sc);
  }
  // rows past M (zero accumulators: the TMA zero-fills them) run through the same arithmetic and only skip their stores,
  // so the warp never diverges here (the early return cost a BSSY / BSYNC pair per chunk: 19 % of the stall samples)
  const bool row_ok = row < g.M;
  float dscale = 1.f;
  if (DROP) {
    dscale = g.drop_inv_keep;
    if (g.sample_scale && row_ok) dscale *= __ldg(g.sample_scale + row / g.rows_per_sample);
  }
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += 16) {
    const int n = n0 + j0;
    if (n >= g.N) break;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j0 + j]);
    if (EPI == O2_EPI_BIAS || EPI == O2_EPI_BIAS_GELU || EPI == O2_EPI_BIAS_RES) {
#pragma unroll
      for (int q = 0; q < 16; q += 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(sbias + j0 + q);
        v[q] += b0.x; v[q + 1] += b0.y; v[q + 2] += b0.z; v[q + 3] += b0.w;
      }
    }
    if (EPI == O2_EPI_BIAS_GELU) {
      if (tma) stage_row16(sc, j0, v);
      else if (row_ok) store_bf16x16(g.aux_out + row * g.ld_aux_out + n, v, n, g.N, g.wide_st);
      if (!(g.dbg & 2)) {
        gelu_pairs<4>(v);
        gelu_pairs<4>(v + 8);
      }
    }
    if (DROP && g.drop == 1) drop_apply16(g, v, row, n, dscale);        // before the residual / GELU' factor
    if (EPI == O2_EPI_BIAS_RES || EPI == O2_EPI_DGELU) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint4 u = ax[(j0 >> 3) + h];
        const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
        const float a[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          if (EPI == O2_EPI_BIAS_RES) { v[8 * h + j] += a[j]; v[8 * h + j + 1] += a[j + 1]; }
          else dgelu_mul2(v[8 * h + j], v[8 * h + j + 1], a[j], a[j + 1]);
        }
      }
    }
    if (DROP && EPI == O2_EPI_BIAS_RES && g.drop == 2) drop_apply16(g, v, row, n, dscale);   // pos_drop(x + pos)
    if (kBf16Out && tma) {
      if (EPI == O2_EPI_BIAS_GELU) {
        keep[j0 >> 3] = pack8(v);
        keep[(j0 >> 3) + 1] = pack8(v + 8);
      } else {
        stage_row16(sc, j0, v);
      }
      continue;
    }
    if (!row_ok) continue;
    if (EPI == O2_EPI_ACCUM) {
      float* c = reinterpret_cast<float*>(g.C) + row * g.ldc + n;
#pragma unroll
      for (int h = 0; h < 16; h += 8)
        if (n + h < g.N) {
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(c + h + j, v[h + j]);
        }
    } else if (C_F32) {
      float* c = reinterpret_cast<float*>(g.C) + row * g.ldc + n;
#pragma unroll
      for (int q = 0; q < 16; q += 4)
        if (n + q < g.N) *reinterpret_cast<float4*>(c + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
    } else {
      store_bf16x16(reinterpret_cast<__nv_bfloat16*>(g.C) + row * g.ldc + n, v, n, g.N, g.wide_st);
    }
  }
  if (tma) {
    if (EPI == O2_EPI_BIAS_GELU) {
      if (!(g.dbg & 1)) {
        stage_store(sc, sc.tm_aux, n0);                // the pre-activation block
        stage_wait(sc);
      }
      stage_row16p(sc, 0, keep[0], keep[1]);
      stage_row16p(sc, 16, keep[2], keep[3]);
    }
    stage_store(sc, sc.tm_c, n0);
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_dispatch(const GemmArgs& g, const uint32_t (&r)[32], long long row, int n0,
                                                  const float* sbias, const uint4 (&ax)[4], const StoreCtx& sc) {
  switch (g.epi) {
    case O2_EPI_NONE:
      if (g.c_f32) epilogue_chunk<BN, O2_EPI_NONE, true>(g, r, row, n0, sbias, ax, sc);
      else epilogue_chunk<BN, O2_EPI_NONE, false>(g, r, row, n0, sbias, ax, sc);
      break;
    case O2_EPI_BIAS:
      if (g.c_f32) epilogue_chunk<BN, O2_EPI_BIAS, true>(g, r, row, n0, sbias, ax, sc);
      else epilogue_chunk<BN, O2_EPI_BIAS, false>(g, r, row, n0, sbias, ax, sc);
      break;
    case O2_EPI_BIAS_GELU:
      if (g.drop) epilogue_chunk<BN, O2_EPI_BIAS_GELU, false, true>(g, r, row, n0, sbias, ax, sc);
      else epilogue_chunk<BN, O2_EPI_BIAS_GELU, false>(g, r, row, n0, sbias, ax, sc);
      break;
    case O2_EPI_BIAS_RES:
      if (g.drop) epilogue_chunk<BN, O2_EPI_BIAS_RES, false, true>(g, r, row, n0, sbias, ax, sc);
      else epilogue_chunk<BN, O2_EPI_BIAS_RES, false>(g, r, row, n0, sbias, ax, sc);
      break;
    case O2_EPI_DGELU:
      if (g.drop) epilogue_chunk<BN, O2_EPI_DGELU, false, true>(g, r, row, n0, sbias, ax, sc);
      else epilogue_chunk<BN, O2_EPI_DGELU, false>(g, r, row, n0, sbias, ax, sc);
      break;
    default: epilogue_chunk<BN, O2_EPI_ACCUM, true>(g, r, row, n0, sbias, ax, sc); break;
  }
}

// MC = 2: CTA pairs (cluster of 2 along M) share the B tile -- each CTA fetches half of it and multicasts it into both
// shared memories, so the L2 -> SM operand traffic per 128 x BN x 64 block drops from 16 + 32 KiB to 16 + 16 KiB (the
// kernel is bound by that traffic: 87 -> 131 FLOP per operand byte at BN = 256).  A stage is refilled only after BOTH
// CTAs' MMAs have drained it (multicast tcgen05.commit onto both "empty" barriers).
template <int BN, int MC, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)   // (a 320-thread block is allocated as 384: 168 registers per thread at most)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_aux, const GemmArgs g) {
  using C = Cfg<BN, PAIR>;
  static_assert(!PAIR || MC == 2, "a CTA pair is a cluster of two");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* sstaging = smem + C::kStages * C::kStageBytes;          // [kEpiWarps][2 KiB], 1 KiB aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sstaging + C::kStagingBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias_all = reinterpret_cast<float*>(sstaging + C::kStagingBytes + 256);   // [kEpiWarps][BN / 2] fp32

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t rank = (MC > 1) ? ptx::cluster_ctarank() : 0u;
  const int first_work = (MC > 1) ? (int)(blockIdx.x / MC) : (int)blockIdx.x;
  const int work_stride = (MC > 1) ? (int)(gridDim.x / MC) : (int)gridDim.x;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if (g.tma_st) {
      ptx::prefetch_tmap(&tmap_c);
      if (g.epi == O2_EPI_BIAS_GELU) ptx::prefetch_tmap(&tmap_aux);
    }
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      // + one arrival per epilogue warp when the bias-gradient side product reads the A tiles of a stage (below)
      ptx::mbar_init(&empty_bar[s], PAIR ? 1 : MC + (g.colsum ? kEpiWarps : 0));   // PAIR: the leader's multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], PAIR ? 2 * kEpiWarps : 32 * kEpiWarps);       // PAIR: one arrival per epilogue warp of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) ptx::tmem_alloc_pair<C::kTmemCols>(tmem_slot);
    else ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (MC > 1) ptx::cluster_sync();       // the peer's barriers are initialised before any multicast can reach them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work items: (split, m block [pair], n block); with MC = 2 the pair (2 mp, 2 mp + 1) shares one item
  const int num_m_items = (g.num_m_blk + MC - 1) / MC;
  const int num_work = num_m_items * g.num_n_blk * g.split_k;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (warp-uniform loop, one elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = first_work; w < num_work; w += work_stride) {
        const int n_blk = w % g.num_n_blk;
        const int rest = w / g.num_n_blk;
        const int m_blk = (rest % num_m_items) * MC + (int)rank;      // may be >= num_m_blk for the odd tail: zero rows
        const int split = rest / num_m_items;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (PAIR) {
            // both CTAs' bytes are counted on the LEADER's barrier (the one its MMA warp waits on); each CTA fills its own
            // shared memory: A rows of its m block, B rows [rank * BN / 2, + BN / 2) of the n block.  No multicast.
            if (ptx::elect_one()) {
              uint8_t* sa = smem + stage * C::kStageBytes;
              uint8_t* sb = sa + kStageA;
              if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
              const uint32_t lbar = ptx::mapa_u32(&full_bar[stage], 0);
              if (!g.a_mn) {
                ptx::tma_load_2d_pair(sa, &tmap_a, lbar, kb * BK, m_blk * BM);
              } else {
#pragma unroll
                for (int i = 0; i < BM / 64; ++i)
                  ptx::tma_load_2d_pair(sa + i * (BK * 128), &tmap_a, lbar, m_blk * BM + i * 64, kb * BK);
              }
              if (!g.b_mn) {
                ptx::tma_load_2d_pair(sb, &tmap_b, lbar, kb * BK, n_blk * BN + (int)rank * (BN / 2));
              } else {
#pragma unroll
                for (int i = 0; i < BN / 128; ++i)
                  ptx::tma_load_2d_pair(sb + i * (BK * 128), &tmap_b, lbar, n_blk * BN + (int)rank * (BN / 2) + i * 64, kb * BK);
              }
            }
            __syncwarp();
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          if (ptx::elect_one()) {
            uint8_t* sa = smem + stage * C::kStageBytes;
            uint8_t* sb = sa + kStageA;
            ptx::mbar_expect_tx(&full_bar[stage], C::kStageBytes);
            if (!g.a_mn) {
              ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                ptx::tma_load_2d(sa + i * (BK * 128), &tmap_a, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
            }
            if (MC == 1) {
              if (!g.b_mn) {
                ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
              } else {
#pragma unroll
                for (int i = 0; i < BN / 64; ++i)
                  ptx::tma_load_2d(sb + i * (BK * 128), &tmap_b, &full_bar[stage], n_blk * BN + i * 64, kb * BK);
              }
            } else {
              constexpr uint16_t kMask = (1u << MC) - 1;
              if (!g.b_mn) {           // rows [rank * BN/MC, +BN/MC) of the B tile, to both CTAs
                ptx::tma_load_2d_mc(sb + rank * (BN / MC) * 128, &tmap_b, &full_bar[stage], kb * BK,
                                    n_blk * BN + (int)rank * (BN / MC), kMask);
              } else {                 // this CTA's share of the 64-column chunks
#pragma unroll
                for (int i = 0; i < BN / 64 / MC; ++i) {
                  const int ci = (int)rank * (BN / 64 / MC) + i;
                  ptx::tma_load_2d_mc(sb + ci * (BK * 128), &tmap_b, &full_bar[stage], n_blk * BN + ci * 64, kb * BK, kMask);
                }
              }
            }
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    // The WHOLE warp runs this loop (warp-uniform control flow: the shared-memory descriptors live in uniform registers and
    // an MMA costs a couple of issue slots); one elected lane executes the tcgen05 instructions.  The first version ran the
    // loop under `if (lane == 0)`: divergent code, descriptors rebuilt in vector registers and moved to uniform ones per MMA --
    // measured in the attention kernels at ~150 cycles per MMA, longer than the 128 cycles a 128 x 256 x 16 MMA takes
    // (ncu, qkv shape: tensor pipe 64 % busy with no memory unit above 48 %).
    if (!PAIR || rank == 0) {             // PAIR: only the leader CTA issues; its MMAs read both CTAs' shared memory
      const uint32_t idesc = ptx::umma_idesc_bf16(PAIR ? 2 * BM : BM, BN, g.a_mn, g.b_mn);
      const uint32_t sbase = ptx::smem_u32(smem);
      const uint64_t da0 = g.a_mn ? ptx::umma_smem_desc(sbase, g.mn_lbo, g.mn_sbo) : ptx::umma_smem_desc(sbase, 16, 1024);
      const uint64_t db0 = g.b_mn ? ptx::umma_smem_desc(sbase + kStageA, g.mn_lbo, g.mn_sbo)
                                  : ptx::umma_smem_desc(sbase + kStageA, 16, 1024);
      // descriptor start-address field counts 16-byte units: advance by (bytes >> 4); k slice of 16: 2048 B (MN-major) / 32 B
      const uint32_t a_step = (g.a_mn ? 2048u : 32u) >> 4, b_step = (g.b_mn ? 2048u : 32u) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = first_work; w < num_work; w += work_stride) {
        const int rest = w / g.num_n_blk;
        const int split = rest / num_m_items;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
        const bool unsummed = g.colsum && (w % g.num_n_blk) != 0;   // the epilogue warps skip this item's main loop
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (PAIR) {
            if (ptx::elect_one()) {
              const uint64_t da = da0 + (uint64_t)((uint32_t)stage * (C::kStageBytes >> 4));
              const uint64_t db = db0 + (uint64_t)((uint32_t)stage * (C::kStageBytes >> 4));
              if (kb == kb0) ptx::umma_pair_first(tmem_d, da, db, idesc);
              else ptx::umma_pair_acc(tmem_d, da, db, idesc);
#pragma unroll
              for (int k = 1; k < BK / 16; ++k) ptx::umma_pair_acc(tmem_d, da + (uint64_t)(k * a_step), db + (uint64_t)(k * b_step), idesc);
              ptx::umma_commit_pair(&empty_bar[stage]);        // frees the stage in BOTH CTAs once these MMAs retire
            }
            __syncwarp();
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          if (ptx::elect_one()) {
            const uint64_t da = da0 + (uint64_t)((uint32_t)stage * (C::kStageBytes >> 4));
            const uint64_t db = db0 + (uint64_t)((uint32_t)stage * (C::kStageBytes >> 4));
            if (kb == kb0) ptx::umma_ss_first(tmem_d, da, db, idesc);
            else ptx::umma_ss_acc(tmem_d, da, db, idesc);
#pragma unroll
            for (int k = 1; k < BK / 16; ++k) ptx::umma_ss_acc(tmem_d, da + (uint64_t)(k * a_step), db + (uint64_t)(k * b_step), idesc);
            if (MC == 1) ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
            else ptx::umma_commit_mc(&empty_bar[stage], (1u << MC) - 1);   // ... in both CTAs of the pair
            if (unsummed) ptx::mbar_arrive_cnt(&empty_bar[stage], kEpiWarps);
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) {                                   // accumulator complete -> epilogue (PAIR: of both CTAs)
          if (PAIR) ptx::umma_commit_pair(&tfull_bar[acc]);
          else ptx::umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------ epilogue: TMEM -> registers -> global
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int cpart = (warp - 2) >> 2;  // which slice of the tile's columns this warp drains
    constexpr int kChunks = (BN / 32) / (kEpiWarps / 4);   // 32-column chunks per warp
    int acc = 0;
    uint32_t acc_phase = 0;
    int cs_stage = 0;                 // the smem ring position of this item's first k block (bias-gradient side product)
    uint32_t cs_phase = 0;
    for (int w = first_work; w < num_work; w += work_stride) {
      const int n_blk = w % g.num_n_blk;
      const int m_blk = ((w / g.num_n_blk) % num_m_items) * MC + (int)rank;
      if (g.colsum) {
        // Bias gradient on the weight-gradient GEMM: dW = dY^T X streams dY through shared memory as the MN-major A
        // operand ([64 k rows][64 m] bf16 boxes, 128B-swizzled); the epilogue warps -- idle during a K loop of
        // T / 64 ~ 2000 blocks -- add up the rows of every A tile of the items with n_blk == 0 (each A tile is loaded once
        // per n block).  A warp reads 4 rows x 128 B per LDS.128 (conflict-free); lane (r4, pc) = (lane >> 3, lane & 7)
        // sees physical 16-byte chunk pc of rows 4 q + r4, i.e. logical chunk pc ^ (row & 7): two accumulator sets by the
        // parity of q.  Warp e of 8: box e >> 2 (m columns 64 box ..), rows 16 (e & 3) .. + 15 of the box.
        const int rest = w / g.num_n_blk;
        const int split = rest / num_m_items;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
        if (n_blk == 0) {
          const int e = warp - 2;
          const int r4 = lane >> 3, pc = lane & 7;
          float cs[2][8];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) cs[i][j] = 0.f;
          // per-item reduction in a FIXED order (no shared-memory atomics): 32 partial rows (4 warps x 4 row lanes x 2
          // parity sets) x 128 columns of fp32 in the TMA-store staging block, which O2_EPI_ACCUM does not use
          float* spart = reinterpret_cast<float*>(sstaging);
          for (int kb = kb0; kb < kb1; ++kb) {
            ptx::mbar_wait(&full_bar[cs_stage], cs_phase);
            const uint8_t* box = smem + cs_stage * C::kStageBytes + (e >> 2) * (BK * 128) + (e & 3) * 16 * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 v = *reinterpret_cast<const uint4*>(box + (4 * q + r4) * 128 + pc * 16);
              const float2 a0 = unpack_bf16x2(v.x), a1 = unpack_bf16x2(v.y), a2 = unpack_bf16x2(v.z), a3 = unpack_bf16x2(v.w);
              float* c = cs[q & 1];
              c[0] += a0.x; c[1] += a0.y; c[2] += a1.x; c[3] += a1.y; c[4] += a2.x; c[5] += a2.y; c[6] += a3.x; c[7] += a3.y;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty_bar[cs_stage]);
            if (++cs_stage == C::kStages) { cs_stage = 0; cs_phase ^= 1; }
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int chunk = pc ^ (4 * i + r4);          // rows 16 (e & 3) + 4 q + r4: (row & 7) = 4 (q & 1) + r4
            const int slot = ((e & 3) * 4 + r4) * 2 + i;
            float* dst = spart + slot * BM + (e >> 2) * 64 + chunk * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(cs[i][0], cs[i][1], cs[i][2], cs[i][3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(cs[i][4], cs[i][5], cs[i][6], cs[i][7]);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
          const int mloc = e * 32 + lane;
          if (mloc < BM && (long long)m_blk * BM + mloc < g.M) {
            float t = 0.f;
#pragma unroll 8
            for (int sl = 0; sl < 32; ++sl) t += spart[sl * BM + mloc];
            atomicAdd(g.colsum + (long long)m_blk * BM + mloc, t);     // one addend per (m block, K split)
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");     // spart may be rewritten by the next item
        } else {
          const int nk = cs_stage + (kb1 - kb0);
          cs_phase ^= (uint32_t)((nk / C::kStages) & 1);
          cs_stage = nk % C::kStages;
        }
      }
      // stage this warp's bias slice (kChunks * 32 columns) in shared memory while the accumulator is still in flight
      const bool has_bias = (g.epi == O2_EPI_BIAS || g.epi == O2_EPI_BIAS_GELU || g.epi == O2_EPI_BIAS_RES);
      const bool has_aux = (g.epi == O2_EPI_BIAS_RES || g.epi == O2_EPI_DGELU);
      const int ncol0 = n_blk * BN + cpart * kChunks * 32;
      float* sb = sbias_all + (warp - 2) * (kChunks * 32);
      __syncwarp();
      if (has_bias) {
        for (int i = lane * 4; i < kChunks * 32; i += 128) {
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ncol0 + i < g.N) bv = *reinterpret_cast<const float4*>(g.bias + ncol0 + i);
          *reinterpret_cast<float4*>(sb + i) = bv;
        }
      }
      __syncwarp();
      const long long row = (long long)m_blk * BM + q * 32 + lane;
      const long long arow = (g.epi == O2_EPI_BIAS_RES) ? (row % g.aux_rows) : row;
      auto load_aux = [&](int c, uint4 (&ax)[4]) {
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const int n = n_blk * BN + c * 32 + j * 8;
          ax[j] = make_uint4(0u, 0u, 0u, 0u);
          ax[j + 1] = make_uint4(0u, 0u, 0u, 0u);
          if (has_aux && row < g.M && n < g.N) {
            const __nv_bfloat16* src = g.aux + arow * g.ld_aux + n;
            if (g.wide_ld && n + 16 <= g.N) {           // one full 32-byte sector per lane and request
              asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                           : "=r"(ax[j].x), "=r"(ax[j].y), "=r"(ax[j].z), "=r"(ax[j].w), "=r"(ax[j + 1].x), "=r"(ax[j + 1].y),
                             "=r"(ax[j + 1].z), "=r"(ax[j + 1].w)
                           : "l"(src));
            } else {
              ax[j] = *reinterpret_cast<const uint4*>(src);
              if (n + 8 < g.N) ax[j + 1] = *reinterpret_cast<const uint4*>(src + 8);
            }
          }
        }
      };
      uint4 ax[4];
      load_aux(cpart * kChunks, ax);                 // first chunk's aux before waiting for the accumulator
      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      // (Prefetching the next chunk's tcgen05.ld under this chunk's math -- two register buffers -- changed nothing: the
      // epilogue was bound by its store path, see StoreCtx.)  The accumulator stage is handed back to the MMA warp as soon as
      // its last chunk is in registers.
      StoreCtx sc;
      sc.stage = g.tma_st ? sstaging + (warp - 2) * 2048 : nullptr;
      sc.tm_c = &tmap_c;
      sc.tm_aux = &tmap_aux;
      sc.row0 = m_blk * BM + q * 32;
      sc.lane = lane;
      const int cbase = cpart * kChunks;
#pragma unroll 1
      for (int i = 0; i < kChunks; ++i) {
        const int c = cbase + i;
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr + c * 32, r);
        uint4 axn[4];
        if (i + 1 < kChunks) load_aux(c + 1, axn);      // next chunk's aux rides under this chunk's math
        ptx::tmem_ld_wait();
#ifndef O2_GEMM_LATE_RELEASE
        if (i + 1 == kChunks) {
          ptx::tc_fence_before();
          if (PAIR) {                                   // one arrival per warp on the LEADER's barrier (its MMA warp waits there)
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(&tempty_bar[acc], 0));
          } else {
            ptx::mbar_arrive(&tempty_bar[acc]);
          }
        }
#endif
        epilogue_dispatch<BN>(g, r, row, n_blk * BN + c * 32, sb + i * 32, ax, sc);
#pragma unroll
        for (int j = 0; j < 4; ++j) ax[j] = axn[j];
      }
#ifdef O2_GEMM_LATE_RELEASE
      ptx::tc_fence_before();
      if (PAIR) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(&tempty_bar[acc], 0));
      } else {
        ptx::mbar_arrive(&tempty_bar[acc]);
      }
#endif
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (warp >= 2 && g.tma_st && lane == 0) ptx::bulk_wait0();   // this lane's TMA stores have left shared memory and landed
  ptx::tc_fence_before();
  __syncthreads();
  if (MC > 1) ptx::cluster_sync();       // no multicast / remote arrive may still target a CTA that has exited
  if (warp == 1) {
    if (PAIR) ptx::tmem_dealloc_pair<C::kTmemCols>(tmem_base);
    else ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BN, int MC, bool PAIR = false>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tx, GemmArgs& g,
           cudaStream_t st) {
  using C = Cfg<BN, PAIR>;
  O2_SET_SMEM_ONCE((gemm_tc_kernel<BN, MC, PAIR>), C::kSmemBytes);
  const int num_work = ((g.num_m_blk + MC - 1) / MC) * g.num_n_blk * g.split_k;
  int ctas = num_work * MC < o2_num_sms() ? num_work * MC : o2_num_sms() / MC * MC;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  O2_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, MC, PAIR>, ta, tb, tc, tx, g));
  return O2_OK;
}

}  // namespace

int o2_gemm_tc(const void* A, int trans_a, int64_t lda, const void* B, int trans_b, int64_t ldb, void* Cp, int c_dtype,
               int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const float* bias, const void* aux,
               int64_t ld_aux, int64_t aux_rows, void* aux_out, int64_t ld_aux_out, int split_k, const O2GemmDrop* drop,
               cudaStream_t st) {
  O2_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_tc: empty problem %lld x %lld x %lld", (long long)M, (long long)N, (long long)K);
  O2_REQUIRE(N % 8 == 0, "gemm_tc: N=%lld must be a multiple of 8", (long long)N);
  O2_REQUIRE((lda * 2) % 16 == 0 && (ldb * 2) % 16 == 0, "gemm_tc: operand row pitch must be 16-byte aligned");
  O2_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)Cp % 16) == 0,
             "gemm_tc: operand base pointers must be 16-byte aligned");
  O2_REQUIRE(ldc % 8 == 0, "gemm_tc: ldc must be a multiple of 8");
  O2_REQUIRE(epilogue >= O2_EPI_NONE && epilogue <= O2_EPI_ACCUM, "gemm_tc: bad epilogue %d", epilogue);
  if (epilogue == O2_EPI_BIAS || epilogue == O2_EPI_BIAS_GELU || epilogue == O2_EPI_BIAS_RES)
    O2_REQUIRE(bias != nullptr, "gemm_tc: epilogue %d needs bias", epilogue);
  if (epilogue == O2_EPI_BIAS_RES || epilogue == O2_EPI_DGELU)
    O2_REQUIRE(aux != nullptr && ld_aux % 8 == 0, "gemm_tc: epilogue %d needs aux (ld multiple of 8)", epilogue);
  if (epilogue == O2_EPI_BIAS_GELU)
    O2_REQUIRE(aux_out != nullptr && ld_aux_out % 8 == 0, "gemm_tc: BIAS_GELU needs aux_out");
  if (epilogue == O2_EPI_ACCUM) O2_REQUIRE(c_dtype == O2_F32, "gemm_tc: ACCUM needs fp32 C");
  if (epilogue == O2_EPI_ACCUM && bias) O2_REQUIRE(trans_a, "gemm_tc: the column-sum side product of ACCUM needs trans_a=1");
  if (epilogue == O2_EPI_BIAS_GELU || epilogue == O2_EPI_BIAS_RES || epilogue == O2_EPI_DGELU)
    O2_REQUIRE(c_dtype == O2_BF16, "gemm_tc: epilogue %d writes bf16", epilogue);
  if (split_k < 1) split_k = 1;
  O2_REQUIRE(split_k == 1 || epilogue == O2_EPI_ACCUM, "gemm_tc: split_k>1 only with O2_EPI_ACCUM");

  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.epi = epilogue; g.c_f32 = (c_dtype == O2_F32);
  g.ldc = ldc; g.ld_aux = ld_aux; g.aux_rows = aux_rows > 0 ? aux_rows : M; g.ld_aux_out = ld_aux_out;
  g.bias = (epilogue == O2_EPI_ACCUM) ? nullptr : bias;
  g.colsum = (epilogue == O2_EPI_ACCUM) ? const_cast<float*>(bias) : nullptr;
  g.aux = (const __nv_bfloat16*)aux; g.aux_out = (__nv_bfloat16*)aux_out; g.C = Cp;
  if (drop) {
    O2_REQUIRE(epilogue == O2_EPI_BIAS_RES || epilogue == O2_EPI_BIAS_GELU || epilogue == O2_EPI_DGELU,
               "gemm_tc: dropout is fused into the BIAS_RES / BIAS_GELU / DGELU epilogues only (got %d)", epilogue);
    O2_REQUIRE(drop->p >= 0.f && drop->p < 1.f, "gemm_tc: dropout p=%f outside [0, 1)", (double)drop->p);
    O2_REQUIRE(!drop->sample_scale || drop->rows_per_sample > 0, "gemm_tc: rows_per_sample must be > 0 with sample_scale");
    O2_REQUIRE(!drop->after_residual || epilogue == O2_EPI_BIAS_RES, "gemm_tc: after_residual is a BIAS_RES option");
    g.drop = drop->after_residual ? 2 : 1;
    g.drop_key = ptx::lowbias32((uint32_t)drop->seed ^ ptx::lowbias32(drop->site ^ (uint32_t)(drop->seed >> 32)));
    g.step_word = o2_step_word();
    g.drop_thr16 = (uint32_t)floor((double)drop->p * 65536.0);
    g.drop_inv_keep = 1.f / (1.f - drop->p);
    g.sample_scale = drop->sample_scale;
    g.rows_per_sample = drop->sample_scale ? drop->rows_per_sample : 1;
  }
  g.wide_st = (ldc % 16 == 0) && ((uintptr_t)Cp % 32 == 0) && (!aux_out || (ld_aux_out % 16 == 0 && (uintptr_t)aux_out % 32 == 0)) &&
              !getenv("O2_GEMM_NARROW_ST");
  g.wide_ld = aux && (ld_aux % 16 == 0) && ((uintptr_t)aux % 32 == 0) && !getenv("O2_GEMM_NARROW_ST");
  g.a_mn = trans_a ? 1 : 0; g.b_mn = trans_b ? 1 : 0;
  g.kb_total = (int)((K + BK - 1) / BK);
  if (split_k > g.kb_total) split_k = g.kb_total;
  g.kb_per_split = (g.kb_total + split_k - 1) / split_k;
  g.split_k = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;
  g.mn_lbo = BK * 128; g.mn_sbo = 1024;
  if (const char* e = getenv("O2_DBG_MN_LBO")) g.mn_lbo = (uint32_t)atoi(e);
  if (const char* e = getenv("O2_DBG_MN_SBO")) g.mn_sbo = (uint32_t)atoi(e);

  const int BN = (N > 128) ? 256 : 128;
  g.num_m_blk = (int)((M + BM - 1) / BM);
  g.num_n_blk = (int)((N + BN - 1) / BN);
  const int MC = (g.num_m_blk >= 2 && !getenv("O2_GEMM_NO_MULTICAST")) ? 2 : 1;
  CUtensorMap ta, tb;
  {
    uint64_t dims[2], str[1];
    uint32_t box[2];
    if (!trans_a) { dims[0] = K; dims[1] = M; box[0] = BK; box[1] = BM; }
    else          { dims[0] = M; dims[1] = K; box[0] = 64; box[1] = BK; }
    str[0] = (uint64_t)lda * 2;
    int rc = o2_make_tmap(&ta, A, 2, 2, dims, str, box, 1);
    if (rc) return rc;
    if (!trans_b) { dims[0] = K; dims[1] = N; box[0] = BK; box[1] = (uint32_t)(BN / MC); }
    else          { dims[0] = N; dims[1] = K; box[0] = 64; box[1] = BK; }
    str[0] = (uint64_t)ldb * 2;
    rc = o2_make_tmap(&tb, B, 2, 2, dims, str, box, 1);
    if (rc) return rc;
  }
  // bf16 outputs leave through TMA stores (32 x 32 boxes, SWIZZLE_64B); fp32 outputs / split-K atomics keep direct stores
  CUtensorMap tc = ta, tx = ta;
#ifdef O2_GEMM_ABLATE
  if (const char* e = getenv("O2_GEMM_DBG")) g.dbg = atoi(e);
#endif
  g.tma_st = (epilogue != O2_EPI_ACCUM && c_dtype == O2_BF16 && !getenv("O2_GEMM_DIRECT_ST")) ? 1 : 0;
  if (g.tma_st) {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[1] = {(uint64_t)ldc * 2};
    uint32_t box[2] = {32, 32};
    int rc = o2_make_tmap(&tc, Cp, 2, 2, dims, str, box, 2);
    if (rc) return rc;
    if (epilogue == O2_EPI_BIAS_GELU) {
      str[0] = (uint64_t)ld_aux_out * 2;
      O2_REQUIRE(((uintptr_t)aux_out % 16) == 0, "gemm_tc: aux_out must be 16-byte aligned");
      rc = o2_make_tmap(&tx, aux_out, 2, 2, dims, str, box, 2);
      if (rc) return rc;
    }
  }
  // cta_group::2 pairs for the forward and data-gradient GEMMs (A K-major): qkv fwd 1325 -> 1427, qkv dgrad 1375 -> 1480,
  // fc2 fwd 1330 -> 1407 TFLOP/s (cuBLASLt: 1400 / 1590 / 1570).  Weight gradients (both operands MN-major) keep the
  // multicast kernel: it is faster for them (fc2 wgrad 1314 vs 1175, D x D wgrad 1250 vs 994 in pair mode) and the epilogue
  // warps of the bias-gradient side product wait on the CTA-local "full" barrier, which in pair mode only the leader has.
  if (MC == 2 && !g.colsum && !g.a_mn && !getenv("O2_GEMM_NO_PAIR"))
    return BN == 256 ? launch<256, 2, true>(ta, tb, tc, tx, g, st) : launch<128, 2, true>(ta, tb, tc, tx, g, st);
  if (MC == 2) return BN == 256 ? launch<256, 2>(ta, tb, tc, tx, g, st) : launch<128, 2>(ta, tb, tc, tx, g, st);
  return BN == 256 ? launch<256, 1>(ta, tb, tc, tx, g, st) : launch<128, 1>(ta, tb, tc, tx, g, st);
}
