// tcgen05 flash attention (bf16) -- placeholder until the kernels land.
#include "common.cuh"
int o2_attn_fwd_tc(const void*, void*, float*, int, int, int, int, float, cudaStream_t) {
  O2_FAIL(O2_ERR_UNSUPPORTED, "attn_fwd_tc: not built yet");
}
int o2_attn_bwd_tc(const void*, const void*, const void*, const float*, void*, float*, int, int, int, int, float,
                   cudaStream_t) {
  O2_FAIL(O2_ERR_UNSUPPORTED, "attn_bwd_tc: not built yet");
}
