// fp32 SIMT GEMM with the same fused epilogues as the tcgen05 path.  This is the exact-parity
// (fp32, <=1e-4 rel) arm of the library: activations, weights and accumulation are all fp32.
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SArgs {
  const float* A; const float* B; float* C;
  long long sa_m, sa_k, sb_n, sb_k, ldc;
  int M, N, K, epi;
  const float* bias; const float* aux; long long ld_aux, aux_rows; float* aux_out; long long ld_aux_out;
  int split_k, k_per_split;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SArgs g) {
  __shared__ float sA[TK][TM + 4];
  __shared__ float sB[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  float acc[4][4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      int m, k;
      if (g.sa_k == 1) { k = i % TK; m = i / TK; } else { m = i % TM; k = i / TM; }
      const int gm = m0 + m, gk = k0 + k;
      sA[k][m] = (gm < g.M && gk < k_end) ? g.A[gm * g.sa_m + gk * g.sa_k] : 0.f;
    }
    for (int i = threadIdx.x; i < TN * TK; i += 256) {
      int n, k;
      if (g.sb_k == 1) { k = i % TK; n = i / TK; } else { n = i % TN; k = i / TN; }
      const int gn = n0 + n, gk = k0 + k;
      sB[k][n] = (gn < g.N && gk < k_end) ? g.B[gn * g.sb_n + gk * g.sb_k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      switch (g.epi) {
        case O2_EPI_BIAS: v += g.bias[n]; break;
        case O2_EPI_BIAS_GELU:
          v += g.bias[n];
          g.aux_out[m * g.ld_aux_out + n] = v;
          v = gelu_f(v);
          break;
        case O2_EPI_BIAS_RES: v += g.bias[n] + g.aux[(m % g.aux_rows) * g.ld_aux + n]; break;
        case O2_EPI_DGELU: v *= dgelu_f(g.aux[m * g.ld_aux + n]); break;
        default: break;
      }
      if (g.epi == O2_EPI_ACCUM) atomicAdd(&g.C[m * g.ldc + n], v);
      else g.C[m * g.ldc + n] = v;
    }
  }
}

}  // namespace

int o2_gemm_simt(const void* A, int trans_a, int64_t lda, const void* B, int trans_b, int64_t ldb, void* C, int c_dtype,
                 int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const float* bias, const void* aux,
                 int64_t ld_aux, int64_t aux_rows, void* aux_out, int64_t ld_aux_out, int split_k, cudaStream_t st) {
  O2_REQUIRE(c_dtype == O2_F32, "gemm_simt: fp32 only");
  O2_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_simt: empty problem");
  O2_REQUIRE(epilogue >= O2_EPI_NONE && epilogue <= O2_EPI_ACCUM, "gemm_simt: bad epilogue %d", epilogue);
  if (split_k < 1) split_k = 1;
  O2_REQUIRE(split_k == 1 || epilogue == O2_EPI_ACCUM, "gemm_simt: split_k>1 only with O2_EPI_ACCUM");
  SArgs g;
  g.A = (const float*)A; g.B = (const float*)B; g.C = (float*)C;
  g.sa_m = trans_a ? 1 : lda; g.sa_k = trans_a ? lda : 1;
  g.sb_n = trans_b ? 1 : ldb; g.sb_k = trans_b ? ldb : 1;
  g.ldc = ldc; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.epi = epilogue;
  g.bias = bias; g.aux = (const float*)aux; g.ld_aux = ld_aux; g.aux_rows = aux_rows > 0 ? aux_rows : M;
  g.aux_out = (float*)aux_out; g.ld_aux_out = ld_aux_out;
  int kps = (int)((K + split_k - 1) / split_k);
  kps = (kps + TK - 1) / TK * TK;
  g.k_per_split = kps;
  g.split_k = (int)((K + kps - 1) / kps);
  dim3 grid((unsigned)((N + TN - 1) / TN), (unsigned)((M + TM - 1) / TM), (unsigned)g.split_k);
  O2_REQUIRE(grid.y <= 65535, "gemm_simt: M too large for this path (%lld)", (long long)M);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(g);
  O2_LAUNCH_CHECK();
  return O2_OK;
}

int o2_gemm_tc(const void* A, int trans_a, int64_t lda, const void* B, int trans_b, int64_t ldb, void* C, int c_dtype,
               int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const float* bias, const void* aux,
               int64_t ld_aux, int64_t aux_rows, void* aux_out, int64_t ld_aux_out, int split_k, const O2GemmDrop* drop,
               cudaStream_t st);

extern "C" int o2_gemm_drop(const void* A, int trans_a, int64_t lda, const void* B, int trans_b, int64_t ldb, void* C,
                            int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const float* bias, const void* aux,
                            int64_t ld_aux, int64_t aux_rows, void* aux_out, int64_t ld_aux_out, const O2GemmDrop* drop,
                            void* stream) {
  O2_REQUIRE(drop != nullptr, "o2_gemm_drop: null dropout descriptor");
  return o2_gemm_tc(A, trans_a, lda, B, trans_b, ldb, C, O2_BF16, ldc, M, N, K, epilogue, bias, aux, ld_aux, aux_rows, aux_out,
                    ld_aux_out, 1, drop, (cudaStream_t)stream);
}

extern "C" int o2_gemm(int impl, const void* A, int trans_a, int64_t lda, const void* B, int trans_b, int64_t ldb,
                       void* C, int c_dtype, int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue,
                       const float* bias, const void* aux, int64_t ld_aux, int64_t aux_rows, void* aux_out,
                       int64_t ld_aux_out, int split_k, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == O2_GEMM_SIMT_F32) {
    const bool side = (epilogue == O2_EPI_ACCUM && bias != nullptr);   // bias gradient next to the weight gradient
    if (side) O2_REQUIRE(trans_a, "o2_gemm: the column-sum side product of O2_EPI_ACCUM needs trans_a=1");
    int rc = o2_gemm_simt(A, trans_a, lda, B, trans_b, ldb, C, c_dtype, ldc, M, N, K, epilogue, side ? nullptr : bias, aux,
                          ld_aux, aux_rows, aux_out, ld_aux_out, split_k, st);
    if (rc || !side) return rc;
    return o2_colsum(A, O2_F32, const_cast<float*>(bias), K, M, lda, stream);   // the fp32 arm keeps the separate pass
  }
  if (impl == O2_GEMM_TC_BF16)
    return o2_gemm_tc(A, trans_a, lda, B, trans_b, ldb, C, c_dtype, ldc, M, N, K, epilogue, bias, aux, ld_aux, aux_rows,
                      aux_out, ld_aux_out, split_k, nullptr, st);
  O2_FAIL(O2_ERR_ARG, "o2_gemm: unknown impl %d", impl);
}
