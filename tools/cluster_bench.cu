// Dev microbenchmark (cluster of 2 CTAs): (1) round-trip latency of remote mbarrier arrives (ping-pong),
// (2) does a 1-D bulk copy issued by CTA 1 into ITS OWN shared memory complete_tx on an mbarrier of CTA 0?
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
__device__ __forceinline__ bool wait(uint32_t bar, uint32_t parity) { int n = 0; while (!try_wait(bar, parity)) { if (++n > (1 << 22)) return false; } return true; }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(bar, rank)) : "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) k(const uint8_t* src, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 32768;       // [0] ping, [8] pong, [16] bulk-full
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + i * 8), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    const int N = 2000;
    // (1) ping-pong
    long long t0 = clock64();
    bool ok = true;
    uint32_t ph = 0;
    for (int i = 0; i < N && ok; ++i) {
      if (rank == 0) { arrive_remote(bar + 0, 1); ok = wait(bar + 8, ph); }
      else { ok = wait(bar + 0, ph); arrive_remote(bar + 8, 0); }
      ph ^= 1u;
    }
    long long t1 = clock64();
    if (rank == 0) { out[0] = (t1 - t0) / N; out[1] = ok; }
    // (2) peer bulk copy signalling the leader's barrier
    if (rank == 1) {
      const uint32_t rbar = mapa(bar + 16, 0);
      const uint32_t dst = mapa(base, 1);
      asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(rbar), "r"(16384) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(src), "r"(16384), "r"(rbar) : "memory");
    } else {
      long long a = clock64();
      bool ok2 = wait(bar + 16, 0);
      out[2] = ok2; out[3] = clock64() - a;
    }
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank == 1 && threadIdx.x == 0) { out[4] = reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[100]; }
}
int main() {
  uint8_t* w; long long* out;
  cudaMalloc(&w, 65536); cudaMemset(w, 0x5a, 65536); cudaMalloc(&out, 64); cudaMemset(out, 0, 64);
  const size_t smem = 1024 + 32768 + 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<2, 64, smem>>>(w, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
  printf("status %s\nping-pong round trip %lld cycles (ok=%lld) -> one-way remote arrive+wake ~%lld\n", cudaGetErrorString(e), h[0], h[1], h[0] / 2);
  printf("peer bulk copy -> leader mbarrier: completed=%lld after %lld cycles, data word 0x%llx (expect 0x5a5a5a5a)\n", h[2], h[3], h[4]);
  return 0;
}
