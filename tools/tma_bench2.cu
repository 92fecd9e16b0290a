// Dev microbenchmark 2: do S bulk copies issued back to back by one thread overlap?  time(batch of S) vs S.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, long long* out, int S, int bytes, int reps, int one_barrier) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 8 * 26624;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + i * 8), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      if (one_barrier) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * S) : "memory");
        for (int i = 0; i < S; ++i)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(base + i * 26624), "l"(src + (size_t)((r * S + i) % 12) * 26624), "r"(bytes), "r"(bar) : "memory");
        int n = 0;
        while (!try_wait(bar, ph) && n < (1 << 22)) ++n;
      } else {
        for (int i = 0; i < S; ++i) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + i * 8), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(base + i * 26624), "l"(src + (size_t)((r * S + i) % 12) * 26624), "r"(bytes), "r"(bar + i * 8) : "memory");
        }
        for (int i = 0; i < S; ++i) { int n = 0; while (!try_wait(bar + i * 8, ph) && n < (1 << 22)) ++n; }
      }
      ph ^= 1u;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}
int main() {
  uint8_t* w; long long* out;
  cudaMalloc(&w, 16 * 26624); cudaMemset(w, 0, 16 * 26624); cudaMalloc(&out, 64);
  const size_t smem = 1024 + 8 * 26624 + 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int grid : {1, 148})
    for (int ob : {0, 1})
      for (int bytes : {2048, 16384, 26624})
        for (int S : {1, 2, 4, 8}) {
          const int reps = 500;
          for (int rep = 0; rep < 2; ++rep) { k<<<grid, 128, smem>>>(w, out, S, bytes, reps, ob); cudaDeviceSynchronize(); }
          long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
          printf("grid=%3d one_barrier=%d bytes=%5d S=%d : %.0f cyc/batch  -> %.1f B/clk/SM\n", grid, ob, bytes, S, (double)h / reps,
                 (double)bytes * S / ((double)h / reps));
        }
  return 0;
}
