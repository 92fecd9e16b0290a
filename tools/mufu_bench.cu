// Dev microbenchmark: issue rate of MUFU flavours on sm_100a (clocks per warp-instruction per SMSP).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  float x[8];
  uint32_t y[8];
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 0.001f + i * 0.1f; y[i] = 0x3f003e80u + threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 2) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(y[i]));
      if (MODE == 3) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(y[i]));
      if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 5) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(y[i]));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(y[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  int iters = 2000;
  k<MODE><<<148, warps_per_sm * 32>>>(out, clk, iters); cudaDeviceSynchronize();
  k<MODE><<<148, warps_per_sm * 32>>>(out, clk, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  double per_warp_instr = (double)h / (iters * 8.0);           // clocks per warp-instr as seen by one warp
  double per_smsp = per_warp_instr / (warps_per_sm / 4.0);      // issue interval per SMSP
  printf("%-22s warps/SM=%2d  clk/warp-instr(seen)=%.2f  interval/SMSP=%.2f  -> %.1f lanes/clk/SM\n", name, warps_per_sm, per_warp_instr, per_smsp, 4 * 32 / per_smsp);
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) { run<0>("tanh.f32", 4); run<1>("ex2.f32", 4); run<2>("tanh.bf16x2", 4); run<3>("ex2.bf16x2", 4); run<4>("rcp.f32", 4); run<5>("tanh.f16x2", 4); }
    if (w == 8) { run<0>("tanh.f32", 8); run<1>("ex2.f32", 8); run<2>("tanh.bf16x2", 8); run<3>("ex2.bf16x2", 8); }
    if (w == 16) { run<0>("tanh.f32", 16); run<1>("ex2.f32", 16); run<2>("tanh.bf16x2", 16); run<3>("ex2.bf16x2", 16); }
  }
  return 0;
}
