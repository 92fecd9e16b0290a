// Dev microbenchmark: cost of per-stage bookkeeping around tcgen05.mma in the issuing thread: tcgen05.commit every
// NM MMAs, tcgen05.fence::after_thread_sync, and an mbarrier try_wait on an already-completed barrier.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t a) {
  uint64_t d = 0; d |= (uint64_t)((a & 0x3FFFFu) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); return ok != 0;
}
// mode bits: 1 = commit per stage (rotating 5 barriers), 2 = fence::after per stage, 4 = try_wait (completed barrier) per stage,
//            8 = different A/B addresses per stage (5-stage ring for A, 7 k-blocks for B), 16 = wait on the commit barrier of 5 stages ago
__global__ void __launch_bounds__(128, 1) k(long long* out, int N, int stages, int NM, int mode) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 5 * 16384;   // A ring 5 x 16 KB, B 7 x 16 KB
  const uint32_t bar = base + 12 * 16384;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(raw + (bar + 128 - smem_u32(raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 12 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + i * 8), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  if (warp == 1 && lane == 0) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar + 7 * 8) : "memory");   // barrier 7: phase 0 complete
    const uint32_t idesc = make_idesc(128, N);
    int st = 0; uint32_t ph = 0;
    long long t0 = clock64();
    for (int s = 0; s < stages; ++s) {
      if (mode & 4) { while (!try_wait(bar + 7 * 8, 0)) {} }
      if ((mode & 16) && s >= 5) { while (!try_wait(bar + st * 8, ph ^ 1u)) {} }
      if (mode & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = (mode & 8) ? a_smem + st * 16384 : a_smem, b0 = (mode & 8) ? b_smem + (s % 7) * 16384 : b_smem;
      for (int ks = 0; ks < NM; ++ks) {
        const uint64_t ad = make_desc_sw128(a0 + ks * 32), bd = make_desc_sw128(b0 + ks * 32);
        const uint32_t acc = (s | ks) ? 1u : 0u;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + (uint32_t)((s / 7) & 3) * 128u * ((mode & 8) ? 1u : 0u)), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
      if (mode & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + st * 8) : "memory");
      if (++st == 5) { st = 0; ph ^= 1u; }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 6 * 8) : "memory");
    int n = 0; while (!try_wait(bar + 6 * 8, 0) && n < (1 << 24)) ++n;
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  const size_t smem = 1024 + 12 * 16384 + 512;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int stages = 2000;
  for (int N : {128, 208})
    for (int NM : {4, 2})
      for (int mode : {0, 1, 2, 4, 7, 8, 15, 31}) {
        for (int rep = 0; rep < 2; ++rep) { k<<<148, 128, smem>>>(out, N, stages, NM, mode); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; } }
        long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("N=%3d MMAs/stage=%d mode=%2d : %.0f cyc/stage (floor %d)\n", N, NM, mode, (double)h / stages, NM * N / 2);
      }
  return 0;
}
