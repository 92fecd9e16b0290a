// Dev microbenchmark: does cluster multicast of 1-D bulk copies raise the per-SM W arrival rate above the chip-wide
// L2->SM cap (~30-36 B/clk/SM when all 148 SMs stream)?  Each CTA of a cluster of CS loads 1/CS of every 16 KB block and
// multicasts it to all CTAs of the cluster; a stage is re-filled once every CTA of the cluster released it.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok; asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); return ok != 0;
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void arrive_remote(uint32_t bar, uint32_t cta) {
  uint32_t ra; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
constexpr int ST = 4, BLK = 16384;
template <int CS>
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, long long* out, int rounds, int nblk) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t ring = base, bar = base + ST * BLK;     // full[ST] at bar, empty[ST] at bar+64
  const uint32_t rank = CS > 1 ? ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + i * 8), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + 64 + i * 8), "r"(CS));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CS > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
  if (threadIdx.x == 0) {
    // loader: each stage expects a full block on the local full barrier; issues own slice to everyone
    int st = 0; uint32_t ph = 0; long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (r >= ST) { int n = 0; while (!try_wait(bar + 64 + st * 8, ph ^ 1u) && n < (1 << 22)) ++n; }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + st * 8), "r"(BLK) : "memory");
      const uint8_t* g = src + (size_t)(r % nblk) * BLK + rank * (BLK / CS);
      if (CS > 1) {
        // peers must have armed their barrier? complete_tx may precede expect_tx (tx-count goes negative), which is legal.
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                     ::"r"(ring + st * BLK + rank * (BLK / CS)), "l"(g), "r"(BLK / CS), "r"(bar + st * 8), "h"((uint16_t)((1u << CS) - 1)) : "memory");
      } else {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + st * BLK), "l"(g), "r"(BLK), "r"(bar + st * 8) : "memory");
      }
      if (++st == ST) { st = 0; ph ^= 1u; }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; }
  } else if (threadIdx.x == 32) {
    // consumer: wait full, release to every CTA of the cluster
    int st = 0; uint32_t ph = 0; long long t0 = clock64(); int bad = 0;
    for (int r = 0; r < rounds; ++r) {
      int n = 0; while (!try_wait(bar + st * 8, ph) && n < (1 << 22)) ++n;
      if (n >= (1 << 22)) { bad = 1; break; }
      if (CS > 1) { for (int c = 0; c < CS; ++c) arrive_remote(bar + 64 + st * 8, c); }
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar + 64 + st * 8) : "memory");
      if (++st == ST) { st = 0; ph ^= 1u; }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) { out[1] = t1 - t0; out[2] = bad; }
  }
  if (CS > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
}
template <int CS> void run(const uint8_t* w, long long* out, int nblk, int grid) {
  const size_t smem = 1024 + ST * BLK + 256;
  cudaFuncSetAttribute(k<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int rounds = 4000;
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(out, 0, 64);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<CS>, w, out, rounds, nblk);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CS=%d err %s\n", CS, cudaGetErrorString(e)); return; }
  }
  long long h[4]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
  printf("cluster %d grid %3d nblk %4d : received %.1f B/clk/SM (consumer), loader %.1f B/clk, timeout=%lld\n", CS, grid, nblk, (double)rounds * BLK / (double)h[1], (double)rounds * BLK / (double)h[0], h[2]);
}
int main() {
  uint8_t* w; long long* out;
  cudaMalloc(&w, 256 * BLK); cudaMemset(w, 0, 256 * BLK); cudaMalloc(&out, 64);
  for (int nblk : {26, 256}) {
    run<1>(w, out, nblk, 148); run<2>(w, out, nblk, 148); run<4>(w, out, nblk, 148); run<8>(w, out, nblk, 144);
  }
  run<1>(w, out, 26, 8); run<2>(w, out, 26, 8);
  return 0;
}
