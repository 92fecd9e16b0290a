// Dev microbenchmark: L2 -> shared memory streaming rate of cp.async.bulk.tensor (TMA) on sm_100a when every
// SM re-reads the same small bf16 matrix (the W_out / W_out^T access pattern of the joint kernels).
//   rows x 64 bf16 boxes (128 B per row, 128B swizzle) through a ring of S stages, no compute: a consumer
//   thread releases a stage as soon as it lands.  Reports bytes/clk/SM and the implied load latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_bench tools/tma_bench.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ int g_waitmode = 0;
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  int n = 0;
  const int mode = g_waitmode;
  if (mode == 0) { while (!try_wait(bar, parity) && n < (1 << 22)) ++n; }
  else if (mode == 1) { while (!test_wait(bar, parity) && n < (1 << 26)) ++n; }
  else { while (!try_wait_hint(bar, parity, 20) && n < (1 << 24)) ++n; }
}

__device__ int g_bulk1d = 0;
__device__ const uint8_t* g_src = nullptr;
template <int CL>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, long long* out, int box_rows,
                                            int nbox_rows_total, int kblocks, int stages, int loads) {
  const int bulk1d = g_bulk1d;
  const uint8_t* src = g_src;
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)box_rows * 128u * CL;
  const uint32_t bar = base + stages * stage_bytes;
  uint32_t rank = 0;
  if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + i * 16), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + i * 16 + 8), "r"(CL));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CL > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const int nrow_blocks = nbox_rows_total / (box_rows * CL);
  if (threadIdx.x == 0) {
    // producer
    long long t0 = clock64();
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < loads; ++i) {
      wait(bar + st * 16 + 8, ph ^ 1u);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + st * 16), "r"(stage_bytes) : "memory");
      const int kb = i % kblocks, rb = (i / kblocks) % nrow_blocks;
      const uint32_t dst = base + st * stage_bytes + rank * (uint32_t)box_rows * 128u;
      const int c0 = kb * 64, c1 = (rb * CL + (int)rank) * box_rows;
      if (bulk1d) {
        const uint8_t* gp = src + ((size_t)(i % 16) * stage_bytes) % (416 * 1024 - stage_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(base + st * stage_bytes), "l"(gp), "r"(stage_bytes), "r"(bar + st * 16) : "memory");
      } else if (CL > 1)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                     ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar + st * 16), "r"(c0), "r"(c1), "h"((uint16_t)((1 << CL) - 1)) : "memory");
      else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar + st * 16), "r"(c0), "r"(c1) : "memory");
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
    (void)t0;
  } else if (threadIdx.x == 32) {
    // consumer: release each stage as soon as it is full (arrive on every CTA of the cluster)
    long long t0 = clock64();
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < loads; ++i) {
      wait(bar + st * 16, ph);
      if (CL > 1) {
        for (uint32_t r = 0; r < (uint32_t)CL; ++r) {
          uint32_t ra;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(bar + st * 16 + 8), "r"(r));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
        }
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar + st * 16 + 8) : "memory");
      }
      if (++st == stages) { st = 0; ph ^= 1u; }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (CL > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMapL2promotion g_promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
template <int CL>
void run(EncodeFn enc, void* w, int rows, int cols, int box_rows, int stages, int grid) {
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, g_promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
  long long* out;
  cudaMalloc(&out, 64);
  const int loads = 2000;
  const size_t smem = 1024 + (size_t)stages * box_rows * 128 * CL + 512;
  cudaFuncSetAttribute(k<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<CL>, map, out, box_rows, rows, cols / 64, stages, loads);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return; }
  }
  long long h = 0;
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / loads;
  const double bytes = (double)box_rows * 128 * CL;
  printf("grid=%3d CL=%d box=%3dx128B (stage %5.1f KB) stages=%d : %.0f cyc/stage  %.1f B/clk/SM  in-flight latency ~%.0f cyc\n", grid, CL,
         box_rows, bytes / 1024, stages, per, bytes / per, per * stages);
  cudaFree(out);
}

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr);
  EncodeFn enc = reinterpret_cast<EncodeFn>(sym);
  const int rows = 416, cols = 512;
  void* w;
  cudaMalloc(&w, (size_t)rows * cols * 2);
  cudaMemset(w, 0, (size_t)rows * cols * 2);
  cudaMemcpyToSymbol(g_src, &w, sizeof(void*));
  CUtensorMapL2promotion promos[3] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
  for (int pi = 0; pi < 3; ++pi) {
    g_promo = promos[pi];
    printf("---- tensor 2D, L2 promotion %d\n", pi);
    for (int stages : {2, 4}) {
      run<1>(enc, w, rows, cols, 208, stages, 148);
      run<1>(enc, w, rows, cols, 52, stages, 148);
    }
  }
  int one = 1;
  cudaMemcpyToSymbol(g_bulk1d, &one, sizeof(int));
  printf("---- 1D bulk copies (contiguous)\n");
  for (int stages : {2, 4, 8}) {
    run<1>(enc, w, rows, cols, 208, stages, 148);
    run<1>(enc, w, rows, cols, 128, stages, 148);
    run<1>(enc, w, rows, cols, 52, stages, 148);
    run<1>(enc, w, rows, cols, 16, stages, 148);
  }
  return 0;
}
