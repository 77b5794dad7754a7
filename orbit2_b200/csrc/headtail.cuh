// shared between headtail.cu (generic kernels, C ABI) and headtail_fast.cu (p = 2, mag = 4 register-tiled kernels)
#pragma once
#include "common.cuh"

namespace o2ht {
constexpr int MAXC = 8;      // output channels / cr
constexpr int MAXCIN = 16;   // C + 4
constexpr int MAXC1 = 64;    // cr * mag^2

struct IdxList { int v[MAXCIN]; };

struct HtArgs {
  const void* head_out; const void* h1; const void* g1; const void* dpreds;   // g1: activated + shuffled branch or NULL
  const float* w_out; const float* b_out; const float* w2; const float* b2;
  void* preds; void* d_head_out; void* dh1;
  float* dw_out; float* db_out; float* dw2; float* db2;
  int B, C, gh, gw, p, mag, cr, Hx, Wx, Ho, Wo, Hs, Ws;   // Hs = Hx*mag, Ws = Wx*mag (extent of the shuffled branch)
};

// fast paths (headtail_fast.cu); return O2_OK, an error code, or kNotApplicable when the shape is outside their domain
constexpr int kNotApplicable = 1;
int headtail_fwd_fast(const HtArgs& a, int dtype, cudaStream_t st);
int headtail_bwd_fast(const HtArgs& a, int dtype, cudaStream_t st);
int conv1_fwd_fast(const float* x, const IdxList& idx, const float* w1, const float* b1, void* h1, void* g1, int dtype, int B,
                   int V, int Hx, int Wx, int cin, int c1, int mag, cudaStream_t st);
int conv1_bwd_fast(const float* x, const IdxList& idx, const void* dh1, float* dw1, float* db1, int dtype, int B, int V,
                   int Hx, int Wx, int cin, int c1, cudaStream_t st);
}  // namespace o2ht
