// Dev probe: tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA) on sm_100a.
//   1. where the accumulator of such an MMA lives in each CTA's tensor memory (expected, from the CuTe 2SM fragment
//      layouts: row m -> lane m, column n < N/2 -> column n; n >= N/2 -> lane m + 64, column n - N/2: a 64 x N
//      accumulator takes 128 lanes x N/2 columns, half the columns of the M = 256 form)
//   2. how the A operand must lie in tensor memory for the TS form (lanes 0-63 only, or duplicated in lanes 64-127)
//   3. the issue-to-completion rate of M = 128 pairs against the N/4-cycle floor (SS and TS)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pair_probe tools/pair_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

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
__device__ __forceinline__ uint32_t pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// MODE 0: SS layout probe; 1: TS, A in lanes 0-63 only; 2: TS, A duplicated in lanes 64-127;
// MODE 3: SS timing; 4: TS timing (A duplicated)
#ifndef SENTINEL
#define SENTINEL 0
#endif
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) k(float* out, long long* tout, int N, int iters) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t a_smem = base;                 // 64 rows x 128 B
  const uint32_t b_smem = base + 16384;         // N/2 rows x 128 B
  const uint32_t bar = base + 16384 + 32768;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(gen + 16384 + 32768 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool timing = MODE >= 3;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(gen)[i] = timing ? 0x3c003c00u : 0u;
  __syncthreads();
  if (!timing) {
    // A[m][0] = m + 1, A[m][1] = 1 (m = 64 rank + r); B[n][0] = 256, B[n][1] = n + 1 (n = N/2 rank + j)
    // K-major SW128: row r at r*128, 16-byte chunk c at position c ^ (r & 7)
    if (threadIdx.x < 64) {
      const int r = threadIdx.x, m = 64 * rank + r;
      *reinterpret_cast<uint32_t*>(gen + r * 128 + ((0 ^ (r & 7)) << 4)) = pack((float)(m + 1), 1.f);
    }
    if (threadIdx.x < N / 2) {
      const int j = threadIdx.x, n = (N / 2) * rank + j;
      *reinterpret_cast<uint32_t*>(gen + 16384 + j * 128 + ((0 ^ (j & 7)) << 4)) = pack(256.f, (float)(n + 1));
    }
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  // zero the accumulator area and (TS) write the A operand: thread = lane of the 128 TMEM lanes (warps 0-3)
  if (warp < 4) {
    const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 512; c += 8) {
      uint32_t z = SENTINEL ? __float_as_uint(7.f) : 0u;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tq + c), "r"(z) : "memory");
    }
    if (MODE == 1 || MODE == 2 || MODE == 4) {
      const int l = warp * 32 + lane;
      const int m = 64 * rank + (l & 63);
      uint32_t w0 = (MODE == 4) ? 0x3c003c00u : ((MODE == 2 || l < 64) ? pack((float)(m + 1), 1.f) : 0u);
      uint32_t wr = (MODE == 4) ? 0x3c003c00u : 0u;
      for (int ks = 0; ks < 4; ++ks)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %2, %2, %2, %2, %2, %2};"
                     ::"r"(tq + 448 + ks * 8), "r"(ks == 0 ? w0 : wr), "r"(wr) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  long long t0 = 0, t1 = 0;
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = make_idesc(128, N);
    const int nit = timing ? iters : 1;
    t0 = clock64();
    for (int it = 0; it < nit; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (!timing && ks > 0) break;
        const uint64_t ad = make_desc_sw128(a_smem + ks * 32), bd = make_desc_sw128(b_smem + ks * 32);
        const uint32_t acc = (it | ks) ? 1u : 0u;
        if (MODE == 0 || MODE == 3)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                       ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                       ::"r"(tmem), "r"(tmem + 448 + ks * 8), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
  }
  {
    int n = 0;
    while (!try_wait(bar, 0) && n < (1 << 24)) ++n;
  }
  if (warp == 1 && lane == 0 && rank == 0) {
    t1 = clock64();
    if (blockIdx.x == 0) tout[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!timing && warp < 4 && blockIdx.x < 2) {
    const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 128; c += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tq + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) out[((size_t)rank * 128 + warp * 32 + lane) * 128 + c + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int MODE>
void probe(const char* name, int N) {
  float* out; long long* tout;
  cudaMalloc(&out, 2 * 128 * 128 * 4); cudaMalloc(&tout, 64);
  cudaMemset(out, 0, 2 * 128 * 128 * 4);
  const size_t smem = 1024 + 16384 + 32768 + 256;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<2, 256, smem>>>(out, tout, N, 1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); return; }
  static float h[2 * 128 * 128];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("== %s (N=%d): accumulator map, entry = (m,n) decoded from D = 256(m+1) + (n+1)\n", name, N);
  int bad = 0, cnt = 0;
  for (int c = 0; c < 2; ++c)
    for (int l = 0; l < 128; ++l)
      for (int col = 0; col < 128; ++col) {
        const float v = h[(c * 128 + l) * 128 + col];
        if (v == 0.f) continue;
        const int iv = (int)v, m = iv / 256 - 1, n = iv % 256 - 1;
        ++cnt;
        // expectation: m = 64 c + (l & 63), n = col + (l >= 64 ? N/2 : 0)
        const bool ok = (m == 64 * c + (l & 63)) && (n == col + (l >= 64 ? N / 2 : 0)) && col < N / 2;
        if (!ok && bad < 12) { printf("  cta %d lane %3d col %3d : value %.0f -> (m=%d, n=%d) UNEXPECTED\n", c, l, col, v, m, n); }
        bad += !ok;
      }
  printf("  %d non-zero entries (expected %d), %d off the expected 2x2 layout\n", cnt, 128 * N, bad);
  for (int c = 0; c < 2; ++c)
    for (int l : {0, 1, 63, 64, 65, 127})
      printf("  cta %d lane %3d: col0 %.0f col1 %.0f col%d %.0f col%d %.0f\n", c, l, h[(c * 128 + l) * 128], h[(c * 128 + l) * 128 + 1],
             N / 2 - 1, h[(c * 128 + l) * 128 + N / 2 - 1], N / 2, h[(c * 128 + l) * 128 + N / 2]);
  cudaFree(out); cudaFree(tout);
}

template <int MODE>
void timing(const char* name, int N) {
  float* out; long long* tout;
  cudaMalloc(&out, 2 * 128 * 128 * 4); cudaMalloc(&tout, 64);
  const int iters = 2000;
  const size_t smem = 1024 + 16384 + 32768 + 256;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE><<<148, 256, smem>>>(out, tout, N, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h = 0;
  cudaMemcpy(&h, tout, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (iters * 4.0);
  printf("%-12s N=%3d : %.1f cyc / MMA (floor N/4 = %.0f) -> %.0f%% of peak\n", name, N, per, N / 4.0, 100.0 * (N / 4.0) / per);
  cudaFree(out); cudaFree(tout);
}

int main() {
#if SENTINEL
  {
    // which columns does an overwriting (accumulate = 0) MMA touch?  every column was preset to 7
    float* out; long long* tout;
    cudaMalloc(&out, 2 * 128 * 128 * 4); cudaMalloc(&tout, 64);
    const size_t smem = 1024 + 16384 + 32768 + 256;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int N : {208, 48, 64}) {
      cudaMemset(out, 0, 2 * 128 * 128 * 4);
      k<0><<<2, 256, smem>>>(out, tout, N, 1);
      cudaDeviceSynchronize();
      static float h[2 * 128 * 128];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      for (int l : {0, 64}) {
        printf("N=%d lane %d cols %d..%d:", N, l, N / 2 - 4, N / 2 + 19);
        for (int c = N / 2 - 4; c < N / 2 + 20; ++c) printf(" %.0f", h[l * 128 + c]);
        printf("\n");
      }
    }
    return 0;
  }
#endif
  probe<0>("SS", 64);
  probe<1>("TS, A in lanes 0-63", 64);
  probe<2>("TS, A duplicated", 64);
  probe<0>("SS", 208);
  for (int N : {64, 128, 208, 256}) { timing<3>("SS M128 cg2", N); timing<4>("TS M128 cg2", N); }
  return 0;
}
