// Dev microbenchmark: issue-to-completion rate of tcgen05.mma kind::f16 (bf16) on sm_100a for
//   mode 0: cta_group::1, A and B from shared memory (SS)
//   mode 1: cta_group::1, A from tensor memory (TS), B from shared memory
//   mode 2: cta_group::2 (CTA pair, M=256), A and B from shared memory
// with M=128 per CTA, N in {128, 208, 256}, K=16 per instruction, operands resident (no TMA), optionally
// with the other warps of the CTA streaming LDS.128 from shared memory to measure interference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(long long* out, int N, int iters, int lds_warps) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base;                 // 128 rows x 128 B  (16 KB, 4 k-steps)
  const uint32_t b_smem = base + 16384;         // 256 rows x 128 B  (32 KB)
  const uint32_t bar = base + 16384 + 32768;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(raw + (bar + 64 - smem_u32(raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if (MODE == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // fill operands with small finite values
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (MODE == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (MODE == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  long long t0 = 0, t1 = 0;
  if (warp == 1 && lane == 0 && (MODE != 2 || rank == 0)) {
    const uint32_t idesc = make_idesc(MODE == 2 ? 256 : 128, N);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = make_desc_sw128(a_smem + ks * 32), bd = make_desc_sw128(b_smem + ks * 32);
        const uint32_t acc = (it | ks) ? 1u : 0u;
        if (MODE == 0)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                       ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        if (MODE == 1)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                       ::"r"(tmem), "r"(tmem + 416 + ks * 8), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        if (MODE == 2)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                       ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    }
    if (MODE == 2)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(bar), "h"((uint16_t)3) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  }
  if (warp >= 2 && warp < 2 + lds_warps) {
    // interference: stream LDS.128 over the operand area while the MMAs run
    uint32_t acc = 0;
    int n = 0;
    while (!try_wait(bar, 0) && n < (1 << 22)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t x, y, z, w;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w)
                     : "r"(base + ((uint32_t)(lane * 16 + i * 512 + warp * 4096) & 0xBFFFu)));
        acc += x + y + z + w;
      }
      ++n;
    }
    if (acc == 0x12345678u) out[100] = acc;
  }
  if (warp == 1) {
    int n = 0;
    while (!try_wait(bar, 0) && n < (1 << 24)) ++n;
    if (lane == 0 && (MODE != 2 || rank == 0)) {
      t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (MODE == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    if (MODE == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int MODE>
void run(const char* name, int N, int lds_warps) {
  long long* out;
  cudaMalloc(&out, 1024);
  cudaMemset(out, 0, 1024);
  const int iters = 2000;
  const size_t smem = 1024 + 16384 + 32768 + 256;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (MODE == 2) ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<MODE>, out, N, iters, lds_warps);
    if (e != cudaSuccess) { printf("%s launch failed: %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h = 0;
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (iters * 4.0);
  const double floor_cyc = 128.0 * N / 256.0;
  printf("%-10s N=%3d lds_warps=%d : %.1f cyc / MMA (floor %.0f) -> %.0f%% of peak\n", name, N, lds_warps, per, floor_cyc,
         100.0 * floor_cyc / per);
  cudaFree(out);
}

int main() {
  for (int lw : {0, 4}) {
    for (int N : {128, 208, 256}) {
      run<0>("SS cta1", N, lw);
      run<1>("TS cta1", N, lw);
      run<2>("SS cta2", N, lw);
    }
  }
  return 0;
}
