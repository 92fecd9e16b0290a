// Dev microbenchmark 3: where does the producer/consumer ring lose time?  1D bulk copies of `bytes` through S stages.
//  V1 two threads (lane 0 of warps 0 and 1), empty/full mbarriers      V2 one thread, issue-ahead by S
//  V3 two full warps (all lanes loop, lane 0 elected for the async ops)  V4 as V1 but consumer arrives from 32 lanes (count 32)
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
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) { int n = 0; while (!try_wait(bar, parity) && n < (1 << 22)) ++n; }
__device__ __forceinline__ void issue(uint32_t dst, const uint8_t* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int V>
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, long long* out, int S, int bytes, int loads) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 8 * 26624;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + i * 16), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + i * 16 + 8), "r"(V == 4 ? 32 : 1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (V == 2) {
    if (threadIdx.x == 0) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < S && i < loads; ++i) issue(base + i * 26624, src + (size_t)(i % 12) * 26624, bytes, bar + i * 16);
      for (int i = 0; i < loads; ++i) {
        wait(bar + st * 16, ph);
        if (i + S < loads) issue(base + st * 26624, src + (size_t)((i + S) % 12) * 26624, bytes, bar + st * 16);
        if (++st == S) { st = 0; ph ^= 1u; }
      }
      if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
  } else if (V == 3) {
    if (warp == 0) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < loads; ++i) {
        wait(bar + st * 16 + 8, ph ^ 1u);
        if (lane == 0) issue(base + st * 26624, src + (size_t)(i % 12) * 26624, bytes, bar + st * 16);
        __syncwarp();
        if (++st == S) { st = 0; ph ^= 1u; }
      }
    } else if (warp == 1) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < loads; ++i) {
        wait(bar + st * 16, ph);
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar + st * 16 + 8) : "memory");
        __syncwarp();
        if (++st == S) { st = 0; ph ^= 1u; }
      }
      if (blockIdx.x == 0 && lane == 0) out[0] = clock64() - t0;
    }
  } else {
    if (threadIdx.x == 0) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < loads; ++i) {
        wait(bar + st * 16 + 8, ph ^ 1u);
        issue(base + st * 26624, src + (size_t)(i % 12) * 26624, bytes, bar + st * 16);
        if (++st == S) { st = 0; ph ^= 1u; }
      }
    } else if (warp == 1 && (V == 4 || lane == 0)) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < loads; ++i) {
        wait(bar + st * 16, ph);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar + st * 16 + 8) : "memory");
        if (++st == S) { st = 0; ph ^= 1u; }
      }
      if (blockIdx.x == 0 && lane == 0) out[0] = clock64() - t0;
    }
  }
}
template <int V> void run(const uint8_t* w, long long* out, int S, int bytes) {
  const size_t smem = 1024 + 8 * 26624 + 256;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int loads = 2000;
  for (int rep = 0; rep < 2; ++rep) { k<V><<<148, 128, smem>>>(w, out, S, bytes, loads); cudaDeviceSynchronize(); }
  long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("V%d S=%d bytes=%5d : %.0f cyc/load  %.1f B/clk/SM\n", V, S, bytes, (double)h / loads, bytes / ((double)h / loads));
}
int main() {
  uint8_t* w; long long* out;
  cudaMalloc(&w, 16 * 26624); cudaMemset(w, 0, 16 * 26624); cudaMalloc(&out, 64);
  for (int bytes : {2048, 16384, 26624})
    for (int S : {2, 4, 8}) {
      run<1>(w, out, S, bytes); run<2>(w, out, S, bytes); run<3>(w, out, S, bytes); run<4>(w, out, S, bytes);
    }
  return 0;
}
