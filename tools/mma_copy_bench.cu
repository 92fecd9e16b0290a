// Dev microbenchmark: how fast can W-like data be brought into shared memory WHILE tcgen05.mma (SS, M=128) is reading
// its operands from shared memory at full rate?  mode 0: 1-D bulk copies by one thread; mode 1: LDG.128 + STS.128 by
// 8 warps.  Reports MMA cycles per instruction and copy bytes/clk/SM, for N = 128 (P3 shape) and N = 208 (P1 shape).
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
__device__ volatile int g_stop;
template <int MODE>
__global__ void __launch_bounds__(384, 1) k(const uint8_t* src, long long* out, int N, int iters, int run_mma, int rot, int nblk) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 16384;     // operands: 16 KB + 32 KB
  const uint32_t ring = base + 49152;                      // 4 x 16 KB copy ring
  const uint32_t bar = ring + 65536;                       // [0] mma done, [8+8i] copy full i
  uint32_t* tptr = reinterpret_cast<uint32_t*>(raw + (bar + 128 - smem_u32(raw)));
  __shared__ volatile int s_done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    s_done = 0;
    for (int i = 0; i < 5; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + i * 8), "r"(1));
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
    long long t0 = clock64();
    if (run_mma) {
      const uint32_t idesc = make_idesc(128, N);
      for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = make_desc_sw128(a_smem + ks * 32), bd = make_desc_sw128(b_smem + ks * 32);
          const uint32_t acc = (it | ks) ? 1u : 0u;
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      int n = 0; while (!try_wait(bar, 0) && n < (1 << 24)) ++n;
    } else {
      while (clock64() - t0 < (long long)iters * 4 * 64) {}
    }
    long long t1 = clock64();
    s_done = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  if (MODE == 0 && warp == 2 && lane == 0) {
    // bulk copies, 4 in flight
    long long t0 = clock64(); long long bytes = 0; int st = 0; uint32_t ph = 0; int issued = rot ? (int)((blockIdx.x * 7u) % (unsigned)nblk) : 0;
    for (int i = 0; i < 4; ++i) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + 8 + i * 8), "r"(16384) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + i * 16384), "l"(src + (size_t)((issued++) % nblk) * 16384), "r"(16384), "r"(bar + 8 + i * 8) : "memory");
    }
    while (!s_done) {
      int n = 0; while (!try_wait(bar + 8 + st * 8, ph) && n < (1 << 22)) ++n;
      bytes += 16384;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + 8 + st * 8), "r"(16384) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + st * 16384), "l"(src + (size_t)((issued++) % nblk) * 16384), "r"(16384), "r"(bar + 8 + st * 8) : "memory");
      if (++st == 4) { st = 0; ph ^= 1u; }
    }
    long long t1 = clock64();
    for (int i = 0; i < 4; ++i) { int n = 0; uint32_t p2 = (i < st) ? ph : ph; while (!try_wait(bar + 8 + ((st + i) % 4) * 8, ((st + i) % 4) < st ? ph : ph) && n < (1 << 22)) ++n; (void)p2; }
    if (blockIdx.x == 0) { out[1] = bytes; out[2] = t1 - t0; }
  }
  if (MODE == 1 && warp >= 4) {
    // LDG.128 + STS.128 by 8 warps (256 threads): 16 KB per round = 4 x 16 B per thread
    const int t = threadIdx.x - 128;
    long long t0 = clock64(); long long bytes = 0; int issued = rot ? (int)((blockIdx.x * 7u) % (unsigned)nblk) : 0;
    while (!s_done) {
      const uint4* g = reinterpret_cast<const uint4*>(src + (size_t)(issued % nblk) * 16384);
      uint4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = __ldg(g + t + i * 256);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ring + (uint32_t)((issued & 3) * 16384 + (t + i * 256) * 16)), "r"(v[i].x), "r"(v[i].y), "r"(v[i].z), "r"(v[i].w) : "memory");
      ++issued; bytes += 16384;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && t == 0) { out[1] = bytes; out[2] = t1 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
template <int MODE> void run(const uint8_t* w, long long* out, int N, int run_mma, int rot, int nblk) {
  const size_t smem = 1024 + 49152 + 65536 + 512;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 4000;
  for (int rep = 0; rep < 2; ++rep) { cudaMemset(out, 0, 64); k<MODE><<<148, 384, smem>>>(w, out, N, iters, run_mma, rot, nblk); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return; } }
  long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
  printf("%s rot=%d nblk=%3d N=%3d mma=%d : %.1f cyc/MMA (floor %d) ; copy %.1f B/clk/SM\n", MODE == 0 ? "bulk   " : "ldg+sts", rot, nblk, N, run_mma, (double)h[0] / (iters * 4.0), N / 2, h[2] ? (double)h[1] / (double)h[2] : 0.0);
}
int main() {
  uint8_t* w; long long* out;
  cudaMalloc(&w, 2048 * 16384); cudaMemset(w, 0, 2048 * 16384); cudaMalloc(&out, 64);
  for (int nblk : {4, 26, 208, 2048}) for (int rot : {0, 1}) for (int m : {0, 1}) run<0>(w, out, 128, m, rot, nblk);
  run<1>(w, out, 128, 1, 1, 26);
  return 0;
}
