// Micro-benchmark for the one-pass attention backward: can the L2 absorb the dQ partial sums?
// Every CTA plays one 128-key tile of one (batch, head): per 128-query tile it fetches Q and dO (2 x 16 KiB, bulk
// async copies) and adds a 128 x 64 fp32 dQ partial (32 KiB) into the accumulator -- the traffic pattern of a fused
// dK/dV/dQ kernel at head dim 64 -- with an optional spin of `delay` cycles standing in for the tensor-core work.
//   mode bit 0: loads      bit 1: reduce through cp.reduce.async.bulk (.add.f32)      bit 2: reduce through red.global.add.v4.f32
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench tools/red_bench.cu ; run: ./red_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
red_kernel(float* acc, const uint8_t* q, const uint8_t* dO, int n_qt, int n_kt, int nbh, int items, int mode, int delay) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sred = reinterpret_cast<float*>(smem);                 // 32 KiB
  uint8_t* sld = smem + 32768;                                  // 2 stages x 32 KiB
  __shared__ uint64_t full[2];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sred[i] = 1.0f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  uint32_t ph[2] = {0, 0};
  for (int it = 0; it < items; ++it) {
    const int w = blockIdx.x + it * gridDim.x;
    const int bh = (w / n_kt) % nbh, kt = w % n_kt;
    const size_t tile_in = (size_t)128 * 64 * 2;                // bf16 tile bytes
    const uint8_t* qb = q + (size_t)bh * n_qt * tile_in;
    const uint8_t* db = dO + (size_t)bh * n_qt * tile_in;
    float* ab = acc + (size_t)bh * n_qt * 8192;
    auto issue_load = [&](int j) {
      const int qt = (kt + j) % n_qt, s = j & 1;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(32768) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(s32(sld + s * 32768)), "l"(qb + (size_t)qt * tile_in), "r"(16384), "r"(s32(&full[s])) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(s32(sld + s * 32768 + 16384)), "l"(db + (size_t)qt * tile_in), "r"(16384), "r"(s32(&full[s])) : "memory");
    };
    if ((mode & 1) && threadIdx.x == 0) { issue_load(0); issue_load(1); }
    for (int j = 0; j < n_qt; ++j) {
      const int qt = (kt + j) % n_qt;
      if (mode & 1) {
        if (threadIdx.x == 0) {
          const int s = j & 1;
          uint32_t ok = 0;
          while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(s32(&full[s])), "r"(ph[s]) : "memory");
          ph[s] ^= 1;
          if (j + 2 < n_qt) issue_load(j + 2);
        }
      }
      if (delay > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < delay) { }
      }
      if (mode & 2) {
        if (threadIdx.x == 0) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                       ::"l"(ab + (size_t)qt * 8192), "r"(s32(sred)), "r"(32768) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (mode & 4) {                                            // thread = one dQ row of 64 floats
        float* row = ab + (size_t)qt * 8192 + threadIdx.x * 64;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(row + 4 * c), "f"(1.0f) : "memory");
      }
      if (mode & 8) {                                            // coalesced variant: warp covers 512 contiguous bytes
        float* base = ab + (size_t)qt * 8192;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(base + (c * 128 + threadIdx.x) * 4), "f"(1.0f) : "memory");
      }
      if (delay > 0) __syncthreads();
    }
    if ((mode & 2) && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
  }
}

int main(int argc, char** argv) {
  const int n_qt = 127, n_kt = 127, nbh = 16;                    // 16 (b,h) pairs resident: 66 MB accumulator, 2 x 33 MB inputs
  int items = argc > 1 ? atoi(argv[1]) : 8;
  float* acc; uint8_t *q, *dO;
  const size_t acc_bytes = (size_t)nbh * n_qt * 32768, in_bytes = (size_t)nbh * n_qt * 16384;
  CK(cudaMalloc(&acc, acc_bytes)); CK(cudaMalloc(&q, in_bytes)); CK(cudaMalloc(&dO, in_bytes));
  CK(cudaMemset(q, 0, in_bytes)); CK(cudaMemset(dO, 0, in_bytes));
  int sms = 0, khz = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const int smem = 3 * 32768 + 1024;
  CK(cudaFuncSetAttribute(red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int modes[] = {1, 2, 3, 4, 8, 5, 9};
  const int delays[] = {0, 1200, 2300};
  printf("SMs %d, max clock %d MHz, %d items/CTA x %d q-tiles, 32 KiB reduce + 32 KiB load per q-tile\n", sms, khz / 1000, items, n_qt);
  for (int mode : modes)
    for (int delay : delays) {
      CK(cudaMemset(acc, 0, acc_bytes));
      red_kernel<<<sms, 128, smem>>>(acc, q, dO, n_qt, n_kt, nbh, 1, mode, delay);      // warm-up
      CK(cudaDeviceSynchronize());
      CK(cudaMemset(acc, 0, acc_bytes));
      CK(cudaEventRecord(e0));
      red_kernel<<<sms, 128, smem>>>(acc, q, dO, n_qt, n_kt, nbh, items, mode, delay);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double iters = (double)items * n_qt;                 // per CTA
      const double bytes = iters * sms * 32768.0;
      // check: every accumulator element must equal the number of partials it received
      double want = (mode & 14) ? (double)items * sms * n_qt * 8192.0 * ((mode & 2 ? 1 : 0) + (mode & 4 ? 1 : 0) + (mode & 8 ? 1 : 0)) : 0.0;
      double got = 0.0;
      if (mode & 14) {
        float* h = (float*)malloc(acc_bytes);
        CK(cudaMemcpy(h, acc, acc_bytes, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < acc_bytes / 4; ++i) got += h[i];
        free(h);
      }
      printf("mode %2d delay %4d: %8.3f ms  %7.1f ns/iter  reduce %6.2f TB/s  load %6.2f TB/s  sum %s\n", mode, delay, ms,
             ms * 1e6 / iters, (mode & 14) ? bytes / ms / 1e9 : 0.0, (mode & 1) ? bytes / ms / 1e9 : 0.0,
             (mode & 14) ? (got == want ? "ok" : "MISMATCH") : "-");
    }
  return 0;
}
