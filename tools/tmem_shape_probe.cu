// Dev probe: which (lane, column) of tensor memory each thread receives for the tcgen05.ld shapes other than 32x32b.
// Lane l, column c of a 128 x 16 block is filled with l * 100 + c through 32x32b stores; warp 0 then reads it back
// with 16x64b.x1, 16x128b.x1 and 16x256b.x1 and prints the registers of every lane.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(int* out) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tptr;
  // every warp fills its own 32 lanes, 16 columns
  {
    uint32_t r[16];
    for (int c = 0; c < 16; ++c) r[c] = (uint32_t)((warp * 32 + lane) * 100 + c);
    const uint32_t a = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                   "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    uint32_t a0, a1, a2, a3;
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(a0) : "r"(tmem));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[lane * 8 + 0] = (int)a0;
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(tmem));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[lane * 8 + 1] = (int)a0; out[lane * 8 + 2] = (int)a1;
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(tmem));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[lane * 8 + 3] = (int)a0; out[lane * 8 + 4] = (int)a1; out[lane * 8 + 5] = (int)a2; out[lane * 8 + 6] = (int)a3;
    // 16x128b.x4 from lane 16 of the quarter: 8 registers
    uint32_t b[8];
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]) : "r"(tmem + (16u << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[256 + lane * 8 + i] = (int)b[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}
int main() {
  int* out; cudaMalloc(&out, 2 * 32 * 8 * 4); cudaMemset(out, 0xff, 2 * 32 * 8 * 4);
  k<<<1, 128>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
  int h[512]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("value = lane*100 + column.  thread : 16x64b.x1 | 16x128b.x1 (2 regs) | 16x256b.x1 (4 regs)\n");
  for (int t = 0; t < 32; ++t)
    printf("t%02d : %5d | %5d %5d | %5d %5d %5d %5d\n", t, h[t * 8], h[t * 8 + 1], h[t * 8 + 2], h[t * 8 + 3], h[t * 8 + 4], h[t * 8 + 5], h[t * 8 + 6]);
  printf("16x128b.x4 at lane offset 16 (8 regs)\n");
  for (int t = 0; t < 32; t += 1) { printf("t%02d :", t); for (int i = 0; i < 8; ++i) printf(" %5d", h[256 + t * 8 + i]); printf("\n"); }
  return 0;
}
