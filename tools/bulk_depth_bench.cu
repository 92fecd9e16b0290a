// Dev microbenchmark: per-SM 1-D bulk copy (global/L2 -> smem) throughput vs bytes in flight (stages x block size),
// for 1 CTA and for all 148 SMs streaming the same W-sized buffer.  Answers: is W streaming latency-bound (Little's law)
// or capped per SM / chip-wide?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); return ok != 0;
}
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, long long* out, int rounds, int nbytes_total, int ST, int BLK, int split) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t ring = base, bar = base + ST * BLK;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + i * 8), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int st = 0; uint32_t ph = 0; long long t0 = clock64(); size_t off = 0;
    const int piece = BLK / split;
    for (int r = 0; r < rounds + ST; ++r) {
      if (r >= ST) { int n = 0; while (!try_wait(bar + st * 8, ph ^ 1u) && n < (1 << 22)) ++n; }
      if (r < rounds) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + st * 8), "r"(BLK) : "memory");
        for (int s = 0; s < split; ++s)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + st * BLK + s * piece), "l"(src + off + s * piece), "r"(piece), "r"(bar + st * 8) : "memory");
        off += BLK; if (off + BLK > (size_t)nbytes_total) off = 0;
      }
      if (++st == ST) { st = 0; ph ^= 1u; }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}
int main() {
  uint8_t* w; long long* out;
  const int total = 425984;
  cudaMalloc(&w, total); cudaMemset(w, 0, total); cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int rounds = 3000;
  for (int grid : {1, 148})
    for (int blk : {4096, 16384, 32768})
      for (int st : {2, 4, 6, 12}) {
        if ((size_t)st * blk > 200 * 1024) continue;
        for (int split : {1, 4}) {
          const size_t smem = 1024 + (size_t)st * blk + 256;
          for (int rep = 0; rep < 2; ++rep) { k<<<grid, 128, smem>>>(w, out, rounds, total, st, blk, split); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; } }
          long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
          printf("grid %3d blk %5d stages %2d split %d (in flight %6d B): %.1f B/clk/SM, %.0f cyc/block\n", grid, blk, st, split, st * blk, (double)rounds * blk / (double)h, (double)h / rounds);
        }
      }
  return 0;
}
