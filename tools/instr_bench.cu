// Dev microbenchmark: per-instruction cost (cycles, one warp, dependent issue) of the bookkeeping instructions that
// surround tcgen05.mma in a role-warp loop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); return ok != 0;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred; asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred)); return pred != 0;
}
__global__ void __launch_bounds__(128, 1) k(long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(raw + (bar + 64 - smem_u32(raw)));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + 8), "r"((1 << 20) - 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");   // phase 0 of bar complete
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  if (warp == 1) {
    const int R = 256;
    long long t[10];
    int acc = 0;
    t[0] = clock64();
    for (int i = 0; i < R; ++i) acc += try_wait(bar, 0);
    t[1] = clock64();
    for (int i = 0; i < R; ++i) acc += test_wait(bar, 0);
    t[2] = clock64();
    for (int i = 0; i < R; ++i) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t[3] = clock64();
    for (int i = 0; i < R; ++i) acc += elect_one();
    t[4] = clock64();
    for (int i = 0; i < R; ++i) __syncwarp();
    t[5] = clock64();
    for (int i = 0; i < R; ++i) { if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8) : "memory"); __syncwarp(); }
    t[6] = clock64();
    for (int i = 0; i < R; ++i) { if (elect_one()) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + 8), "r"(0) : "memory"); __syncwarp(); }
    t[7] = clock64();
    for (int i = 0; i < R; ++i) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    t[8] = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { for (int i = 0; i < 8; ++i) out[i] = (t[i + 1] - t[i]); out[9] = acc; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
  long long* out; cudaMalloc(&out, 128);
  for (int rep = 0; rep < 2; ++rep) { k<<<148, 128, 4096>>>(out); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; } }
  long long h[10]; cudaMemcpy(h, out, 80, cudaMemcpyDeviceToHost);
  const char* names[8] = {"try_wait (complete)", "test_wait (complete)", "tcgen05.fence::after", "elect.sync", "__syncwarp", "elect+commit+syncwarp", "elect+arrive.expect_tx+syncwarp", "fence.proxy.async.shared"};
  for (int i = 0; i < 8; ++i) printf("%-34s %.1f cycles each\n", names[i], (double)h[i] / 256.0);
  return 0;
}
