// Shared device/host helpers for libo2b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/o2b200.h"

// ---------------------------------------------------------------- error plumbing
void o2_set_error(const char* fmt, ...);
const uint64_t* o2_step_word();     // abi.cu: the calling thread's current o2_dropout_seed_source pointer (or NULL)
#define O2_FAIL(code, ...)            \
  do {                                \
    o2_set_error(__VA_ARGS__);        \
    return (code);                    \
  } while (0)
#define O2_REQUIRE(cond, ...)                         \
  do {                                                \
    if (!(cond)) O2_FAIL(O2_ERR_ARG, __VA_ARGS__);    \
  } while (0)
#define O2_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess)                                                                \
      O2_FAIL(O2_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)
#define O2_LAUNCH_CHECK() O2_CUDA(cudaGetLastError())
// opt a kernel into > 48 KiB of dynamic shared memory exactly once per process, safely from concurrent host threads
// (the header promises re-entrant entry points; one process drives one GPU)
#define O2_SET_SMEM_ONCE(kernel, bytes)                                                                     \
  do {                                                                                                      \
    static std::once_flag once__;                                                                           \
    static cudaError_t err__ = cudaSuccess;                                                                 \
    std::call_once(once__, [&] {                                                                            \
      err__ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));      \
    });                                                                                                     \
    if (err__ != cudaSuccess)                                                                               \
      O2_FAIL(O2_ERR_CUDA, "%s:%d cudaFuncSetAttribute(%s) -> %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(err__)); \
  } while (0)

static inline int o2_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---------------------------------------------------------------- scalar conversion
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// exact (erf) GELU and derivative -- torch.nn.GELU() default, mlp.py:64 / res_slimvit.py:109,118
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Branch-free erf for the bf16 (tcgen05) epilogues: Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 -- two MUFU ops
// (rcp, ex2) and ~10 FMA-pipe ops instead of libdevice's erff.  The fp32 parity arm keeps erff (gelu_f / dgelu_f).
__device__ __forceinline__ float erf_fast_abs(float ax, float& e_x2) {   // ax = |x|; returns erf(|x|), e_x2 = exp(-x^2)
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));
  e_x2 = e;
  return fmaf(-p, e, 1.0f);
}
__device__ __forceinline__ float gelu_fast(float x) {
  float e;
  const float er = copysignf(erf_fast_abs(fabsf(x) * 0.70710678118654752f, e), x);
  return 0.5f * x * (1.0f + er);
}
__device__ __forceinline__ float dgelu_fast(float x) {
  float e;   // e = exp(-x^2 / 2)
  const float er = copysignf(erf_fast_abs(fabsf(x) * 0.70710678118654752f, e), x);
  return fmaf(x * 0.39894228040143268f, e, 0.5f * (1.0f + er));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- sm_100a PTX wrappers
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (-> cudaErrorLaunchFailure) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("o2b200: mbarrier timeout block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// ---- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA (cp.async.bulk.tensor), completion on an mbarrier
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA reduce: the shared-memory box is ADDED (fp32, element-wise, L2 atomics) into the global tensor at the given
// coordinates; elements outside the tensor are dropped.  Completion is tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void tma_reduce_add_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk groups have finished READING their shared-memory source (it may be overwritten)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... and have completed entirely (global writes performed)
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// multicast: the tile lands at the same shared-memory offset of every CTA in cta_mask and completes tx bytes on the
// mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- TMEM
template <int kCols> __device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int kCols> __device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols));
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 in, fp32 accumulate)
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A is M x K (K-major, 16-bit elements packed two per 32-bit column)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accum)
      : "memory");
}
// variants with the accumulate flag fixed at compile time (no predicate set-up in the issue path)
__device__ __forceinline__ void umma_ss_first(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_ss_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_ts_acc(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// same, arriving on the barrier at this offset in every CTA of cta_mask (stage release across a multicast pair)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
// ---- CTA pairs (cta_group::2): one tcgen05.mma issued by the leader CTA (cluster rank 0) computes a 256-row tile, each CTA
// supplying its own 128 rows of A and HALF of the B tile from its own shared memory (same offsets in both CTAs) and
// receiving its 128 accumulator rows in its own TMEM.  Shared-memory operand traffic per SM: A + B/2 instead of A + B.
template <int kCols> __device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
template <int kCols> __device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols));
}
__device__ __forceinline__ void umma_pair_first(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_pair_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc)
      : "memory");
}
// arrive (once all previously issued MMAs of this thread are complete) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
// shared::cluster address of `p`'s offset in the CTA of cluster rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose completion bytes are counted on a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// store 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// ---- attention-probability dropout (components/attention.py:75 attn_drop), bit-sliced: ONE 32-bit keep word serves the
// 32 keys [32 kb, 32 kb + 32) of one query row q of one (batch, head):
//     base   = lowbias32((q * nkb + kb) ^ key_bh)                       nkb = ceil(N / 32)
//     w_i    = lo32(base * K_i) ^ hi32(base * K_i)    i = 0..7         (eight bit planes of a uniform byte U per key)
//     keep   = (U >= thr), evaluated on all 32 lanes of the planes at once, LSB plane first:
//              ge = ~0;  ge = thr bit i ? (w_i & ge) : (w_i | ge)        -> one LOP3 per plane (tm[i] = bit i ? ~0 : 0)
//     key kk = k & 31 reads bit 7 - (kk >> 2) + 8 (kk & 1) + 16 ((kk >> 1) & 1): the byte-msb order that lets PRMT's
//     sign-replicate mode expand four decisions of (word << s) into bf16x2 / fp32 AND-masks.
// The drop probability has 16-bit resolution (like the token-stream dropout of dropout.cu): thr16 = floor(p * 65536) =
// 256 hi8 + frac8, and the byte threshold of a WORD is hi8 + 1 with probability frac8 / 256 (dither byte
// d = lowbias32(base ^ 0x68E31DA4) >> 24, d < frac8) and hi8 otherwise, so that P(drop) = thr16 / 65536 exactly
// (p = 0.1 -> 0.099991, not 25 / 256 = 0.0977); both comparators are evaluated on the planes (8 more LOP3 per 32 keys).
// key_bh = lowbias32(site_key ^ (b * heads + h) * 0x9E3779B1); kept values are scaled by 65536 / (65536 - thr16), the
// exact keep probability (oracle/dropout_mask.py restates this for the parity tests).
// Row-owner kernels (forward, dQ: thread = query row) build one word per 32 keys; the dK/dV kernels (thread = key row)
// let lane l build the word of query q0 + l and read the others' by shuffle.
struct AttnDrop {
  uint32_t site_key, thr16, nkb, frac8;
  const uint64_t* step_word;  // o2_dropout_seed_source: device word XORed into the seed at run time (CUDA-graph replay), or NULL
  float inv_keep;
  uint32_t tm[8], tn[8];      // plane masks of the byte thresholds hi8 and hi8 + 1
};
__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x21f0aaadu;
  x ^= x >> 15; x *= 0x735a2d97u;
  x ^= x >> 15;
  return x;
}
// contribution of the device-resident step word to every dropout key (0 when none is installed): the masks of a captured
// step change between graph replays because the host rewrites the word, not the kernel arguments
__device__ __forceinline__ uint32_t step_word_mix(const uint64_t* step_word) {
  if (!step_word) return 0u;
  const uint64_t w = *step_word;
  return lowbias32((uint32_t)w ^ lowbias32((uint32_t)(w >> 32) ^ 0x5BD1E995u));
}
__device__ __forceinline__ uint32_t attn_drop_key(const AttnDrop& d, int bh) {
  return lowbias32((d.site_key ^ step_word_mix(d.step_word)) ^ ((uint32_t)bh * 0x9E3779B1u));
}
__device__ __forceinline__ uint32_t attn_keep_word(const AttnDrop& d, uint32_t key_bh, uint32_t q, uint32_t kb) {
  constexpr uint32_t kMul[8] = {0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du, 0x27D4EB2Fu,
                                0x165667B1u, 0xD3A2646Du, 0xFD7046C5u, 0xB55A4F09u};
  const uint32_t base = lowbias32((q * d.nkb + kb) ^ key_bh);
  uint32_t ge = 0xFFFFFFFFu, gh = 0xFFFFFFFFu;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint64_t m = (uint64_t)base * kMul[i];
    const uint32_t w = (uint32_t)m ^ (uint32_t)(m >> 32);
    ge = (w & ge) | (~d.tm[i] & (w | ge));
    gh = (w & gh) | (~d.tn[i] & (w | gh));
  }
  return ((lowbias32(base ^ 0x68E31DA4u) >> 24) < d.frac8) ? gh : ge;
}
// bit of the keep word that belongs to key k (see above)
__host__ __device__ __forceinline__ uint32_t attn_keep_bit(uint32_t k) {
  const uint32_t kk = k & 31u;
  return 7u - (kk >> 2) + 8u * (kk & 1u) + 16u * ((kk >> 1) & 1u);
}
// inverse of attn_keep_bit: the key (0..31) whose decision sits at bit b
__host__ __device__ __forceinline__ uint32_t attn_keep_bit_inv(uint32_t b) {
  return 4u * (7u - (b & 7u)) + 2u * (b >> 4) + ((b >> 3) & 1u);
}
// 32 x 32 bit-matrix transpose across the lanes of a warp: lane l passes row l, lane j gets column j (bit c of the result =
// bit j of lane c's word).  Five butterfly exchanges (Hacker's Delight 7-3 with the row dimension spread over the lanes).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
    x = (lane & s) ? ((x & ~m) | ((y >> s) & m)) : ((x & m) | ((y << s) & ~m));
  }
  return x;
}
// AND-masks of keys (4 s + 2 t, 4 s + 2 t + 1) of a keep word, given ws = word << s: a bf16x2 mask, or two fp32 masks
template <int T2> __device__ __forceinline__ uint32_t keep_mask_bf16x2(uint32_t ws) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(m) : "r"(ws), "n"(T2 ? 0xBBAA : 0x9988));
  return m;
}
template <int T2, int E> __device__ __forceinline__ uint32_t keep_mask_f32(uint32_t ws) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(m) : "r"(ws), "n"(T2 ? (E ? 0xBBBB : 0xAAAA) : (E ? 0x9999 : 0x8888)));
  return m;
}
inline AttnDrop make_attn_drop(float p, uint64_t seed, uint32_t site, int N) {
  AttnDrop d;
  d.step_word = ::o2_step_word();
  d.site_key = lowbias32((uint32_t)seed ^ lowbias32(site ^ (uint32_t)(seed >> 32)));
  double pc = (double)p;
  if (pc > 0.99) pc = 0.99;                                   // hi8 + 1 must stay a byte
  d.thr16 = pc > 0.0 ? (uint32_t)floor(pc * 65536.0) : 0u;
  d.frac8 = d.thr16 & 0xFFu;
  const uint32_t hi8 = d.thr16 >> 8;
  d.nkb = (uint32_t)((N + 31) >> 5);
  d.inv_keep = 65536.f / (65536.f - (float)d.thr16);
  for (int i = 0; i < 8; ++i) {
    d.tm[i] = ((hi8 >> i) & 1u) ? 0xFFFFFFFFu : 0u;
    d.tn[i] = (((hi8 + 1u) >> i) & 1u) ? 0xFFFFFFFFu : 0u;
  }
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ---- packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2: two fp32 lanes per issue slot) and exp2 on the FMA pipe
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// 2^x for a pair, entirely on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + r with the 1.5*2^23 magic
// constant, degree-3 minimax polynomial for 2^r on [-0.5, 0.5] (max relative error 1.3e-4, below bf16 resolution), exponent
// insertion with one LEA.  Inputs are clamped at -125 (result ~2^-125 instead of 0); valid for x < 128.
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t X) {
  float x0, x1;
  unpack2(X, x0, x1);
  X = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const uint64_t T = add2(X, 0x4B4000004B400000ull);                 // x + 12582912
  uint64_t R = add2(T, 0xCB400000CB400000ull);                       // n (as float)
  R = fma2(R, 0xBF800000BF800000ull, X);                             // r = x - n
  uint64_t P = fma2(R, pack2(0.0553272026f, 0.0553272026f), pack2(0.242996181f, 0.242996181f));
  P = fma2(P, R, pack2(0.693246777f, 0.693246777f));
  P = fma2(P, R, pack2(0.999868923f, 0.999868923f));
  uint32_t p0, p1, t0, t1;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(P));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(t0), "=r"(t1) : "l"(T));
  return pack2u(p0 + (t0 << 23), p1 + (t1 << 23));
}
// exp2 of a pair: MUFU for both lanes, or the polynomial when kPoly
template <bool kPoly> __device__ __forceinline__ uint64_t exp2_pair(uint64_t X) {
  if (kPoly) return exp2_poly2(X);
  float x0, x1;
  unpack2(X, x0, x1);
  return pack2(ex2(x0), ex2(x1));
}
__device__ __forceinline__ uint32_t pack_bf16x2_pair(uint64_t v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return pack_bf16x2(lo, hi);
}

__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t cnt) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(cnt) : "memory");
}

// ---- UMMA descriptors (cute/arch/mma_sm100_desc.hpp field layout)
// shared-memory matrix descriptor, SWIZZLE_128B, version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;   // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor for kind::f16, bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)            // c_format = F32
         | (1u << 7)          // a_format = BF16
         | (1u << 10)         // b_format = BF16
         | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

}  // namespace ptx

// ---------------------------------------------------------------- TMA tensor-map creation (host)
// cuTensorMapEncodeTiled is fetched with cudaGetDriverEntryPoint so the library has no link-time
// dependency on libcuda (it must dlopen on the GPU-less build box for the symbol-export test).
int o2_make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes /*rank-1 entries, dims[1..]*/, const uint32_t* box,
                 int swizzle128 /*0 none, 1 SWIZZLE_128B, 2 SWIZZLE_64B*/);
