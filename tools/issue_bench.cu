// Dev microbenchmark: instruction overhead of the tcgen05.mma issuing thread for a realistic stage loop (5-stage A ring,
// 7 B k-blocks, commit per stage, wait on the barrier committed 5 stages earlier).  Variants:
//   0: `if (lane == 0)` single-thread loop, descriptors rebuilt per MMA (the style the kernels used)
//   1: whole warp runs the loop, elect.sync guards MMAs + commit
//   2: as 1, descriptor low words precomputed, 64-bit adds only
//   3: as 0 with precomputed descriptors
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
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
template <int VAR>
__global__ void __launch_bounds__(128, 1) k(long long* out, int nmb, int KBG) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 5 * 16384;
  const uint32_t bar = base + 12 * 16384;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(raw + (bar + 128 - smem_u32(raw)));
  const int warp = (VAR >= 4) ? __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0) : (int)(threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
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
  const uint32_t idesc = make_idesc(128, 128);
  if (warp == 1) {
    long long t0 = clock64();
    if (VAR == 0 || VAR == 3) {
      if (lane == 0) {
        int st = 0; uint32_t ph = 0; int s = 0;
        for (int mb = 0; mb < nmb; ++mb)
          for (int kb = 0; kb < KBG; ++kb, ++s) {
            if (s >= 5) { while (!try_wait(bar + st * 8, ph ^ 1u)) {} }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (VAR == 0) {
              for (int ks = 0; ks < 4; ++ks)
                mma(tmem + (mb & 3) * 128, make_desc_sw128(a_smem + st * 16384 + ks * 32), make_desc_sw128(b_smem + kb * 16384 + ks * 32), idesc, (kb | ks) ? 1u : 0u);
            } else {
              const uint64_t ad = make_desc_sw128(a_smem) + (uint64_t)(st * 1024), bd = make_desc_sw128(b_smem) + (uint64_t)(kb * 1024);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) mma(tmem + (mb & 3) * 128, ad + 2 * ks, bd + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
            }
            commit(bar + st * 8);
            if (++st == 5) { st = 0; ph ^= 1u; }
          }
        commit(bar + 6 * 8);
        int n = 0; while (!try_wait(bar + 6 * 8, 0) && n < (1 << 24)) ++n;
      }
    } else {
      int st = 0; uint32_t ph = 0; int s = 0;
      const uint64_t ad0 = make_desc_sw128(a_smem), bd0 = make_desc_sw128(b_smem);
      for (int mb = 0; mb < nmb; ++mb)
        for (int kb = 0; kb < KBG; ++kb, ++s) {
          if (s >= 5) { while (!try_wait(bar + st * 8, ph ^ 1u)) {} }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            if (VAR == 1 || VAR == 5) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                mma(tmem + (mb & 3) * 128, make_desc_sw128(a_smem + st * 16384 + ks * 32), make_desc_sw128(b_smem + kb * 16384 + ks * 32), idesc, (kb | ks) ? 1u : 0u);
            } else {
              const uint64_t ad = ad0 + (uint64_t)(st * 1024), bd = bd0 + (uint64_t)(kb * 1024);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) mma(tmem + (mb & 3) * 128, ad + 2 * ks, bd + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
            }
            commit(bar + st * 8);
          }
          __syncwarp();
          if (++st == 5) { st = 0; ph ^= 1u; }
        }
      if (elect_one()) commit(bar + 6 * 8);
      __syncwarp();
      int n = 0; while (!try_wait(bar + 6 * 8, 0) && n < (1 << 24)) ++n;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
template <int VAR> void run(long long* out) {
  const size_t smem = 1024 + 12 * 16384 + 512;
  cudaFuncSetAttribute(k<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int nmb = 400, KBG = 7;
  for (int rep = 0; rep < 2; ++rep) { k<VAR><<<148, 128, smem>>>(out, nmb, KBG); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return; } }
  long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("variant %d : %.0f cyc/stage (floor 256)\n", VAR, (double)h / (nmb * KBG));
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  run<0>(out); run<1>(out); run<2>(out); run<3>(out); run<4>(out); run<5>(out);
  return 0;
}
