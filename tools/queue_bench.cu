// Dev microbenchmark: depth of the tcgen05.mma issue queue.  One elected lane issues K MMAs (M=128, N=128, K=16: 64 cycles
// each) back to back; t_issue = cycles until the K-th issue returned, t_done = cycles until a commit after them completed.
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred; asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred)); return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
template <int K>
__global__ void __launch_bounds__(128, 1) k(long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 32768;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(raw + (bar + 64 - smem_u32(raw)));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  if (warp == 1) {
    const uint32_t idesc = make_idesc(128, 128);
    const uint64_t ad = make_desc_sw128(base), bd = make_desc_sw128(base + 16384);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
#pragma unroll
      for (int i = 0; i < K; ++i) mma(tmem, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, i ? 1u : 0u);
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      int n = 0; while (!try_wait(bar, 0) && n < (1 << 24)) ++n;
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
template <int K> void run(long long* out) {
  const size_t smem = 1024 + 32768 + 256;
  cudaFuncSetAttribute(k<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 3; ++rep) { k<K><<<148, 128, smem>>>(out); cudaDeviceSynchronize(); }
  long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("K=%2d MMAs: issue returned after %4lld cycles, all done after %4lld (exec floor %d)\n", K, h[0], h[1], K * 64);
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  run<1>(out); run<2>(out); run<4>(out); run<8>(out); run<16>(out); run<32>(out); run<64>(out);
  return 0;
}
