// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate, M = 128, K = 16) as a function of N
// and of where the A operand lives (shared memory "SS" / tensor memory "TS"), one CTA per SM issuing a long back-to-back
// stream from one elected thread.  It answers the question behind the attention kernels at head dim 64:
// three of the five backward GEMMs (dV, dK, dQ) and P V of the forward have N = hd = 64 -- does an N = 64 instruction
// run at the math rate (32 cycles) or does it take as long as an N = 128 one (64 cycles)?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate tools/mma_rate.cu ; run: ./mma_rate
// Operand contents are irrelevant (zeros); descriptors are the K-major SWIZZLE_128B ones the kernels use.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* cycles, int n_mma) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    const uint64_t da = smem_desc(s32(smem), 16, 1024);              // A: 128 rows x 64 (K-major)
    const uint64_t db = smem_desc(s32(smem) + 16384, 16, 1024);      // B: up to 256 rows x 64
    const uint32_t id = idesc_bf16(128, N);
    const uint32_t d = tm, a_tm = tm + 256;                          // accumulator columns [0, N), A (TS) at column 256
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t k = (uint32_t)(i & 3);
      if (TS) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                     ::"r"(d), "r"(a_tm + k * 8), "l"(db + (uint64_t)(k * 2)), "r"(id) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(d), "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(id) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(s32(&bar)) : "memory");
    }
    const long long t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int N, bool TS>
void run(const char* name, long long* dcyc, int sms) {
  const int n_mma = 4096;
  const size_t smem = 16384 + 32768 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rate_kernel<N, TS><<<sms, 128, smem>>>(dcyc, n_mma);             // warm-up
  rate_kernel<N, TS><<<sms, 128, smem>>>(dcyc, n_mma);
  CK(cudaDeviceSynchronize());
  long long* h = (long long*)malloc(sizeof(long long) * sms);
  CK(cudaMemcpy(h, dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  double s = 0;
  for (int i = 0; i < sms; ++i) s += (double)h[i];
  const double per = s / sms / n_mma;
  const double math = 128.0 * N * 16 * 2 / 8192.0;                   // cycles at 8192 dense bf16 FLOP per clock and SM
  printf("%-10s N = %3d : %6.1f cycles per MMA (math alone: %5.1f)  -> %5.1f %% of the tensor-pipe rate\n", name, N, per, math,
         100.0 * math / per);
  free(h);
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long* d;
  CK(cudaMalloc(&d, sizeof(long long) * sms));
  printf("tcgen05.mma kind::f16, M = 128, K = 16, %d SMs, 4096 back-to-back instructions per SM\n", sms);
  run<32, false>("SS", d, sms);
  run<64, false>("SS", d, sms);
  run<128, false>("SS", d, sms);
  run<256, false>("SS", d, sms);
  run<32, true>("TS", d, sms);
  run<64, true>("TS", d, sms);
  run<128, true>("TS", d, sms);
  run<256, true>("TS", d, sms);
  return 0;
}
