// Microbenchmark: MUFU.EX2 throughput for f32 vs packed f16x2 / bf16x2, and F2FP pack rate.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
template <int MODE>
__global__ void rate(int iters, unsigned long long* cycles, uint32_t* sink, uint32_t seed) {
  uint32_t x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = seed + k * 0x00010001u + threadIdx.x;
  float fa = __uint_as_float(0x3f000000u + seed), fb = fa * 0.5f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[k]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[k]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[k]));
      if (MODE == 3) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(x[k]) : "f"(fa), "f"(fb));
      if (MODE == 4) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x[k]) : "f"(fa), "f"(fb));
      if (MODE == 5) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(x[k]) : "r"(seed));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t s = 0;
  for (int k = 0; k < 16; ++k) s ^= x[k];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (s == 0x12345u) sink[0] = s;
}
int main() {
  unsigned long long* cyc; uint32_t* sink;
  cudaMallocManaged(&cyc, 148 * 8); cudaMalloc(&sink, 64);
  const int iters = 2000;
  const char* names[6] = {"ex2.f32", "ex2.f16x2", "ex2.bf16x2", "cvt.bf16x2.f32", "cvt.f16x2.f32", "add.f16x2"};
  for (int threads : {256, 512}) {
    for (int m = 0; m < 6; ++m) {
      switch (m) {
        case 0: rate<0><<<148, threads>>>(iters, cyc, sink, 1); break;
        case 1: rate<1><<<148, threads>>>(iters, cyc, sink, 1); break;
        case 2: rate<2><<<148, threads>>>(iters, cyc, sink, 1); break;
        case 3: rate<3><<<148, threads>>>(iters, cyc, sink, 1); break;
        case 4: rate<4><<<148, threads>>>(iters, cyc, sink, 1); break;
        default: rate<5><<<148, threads>>>(iters, cyc, sink, 1); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s failed: %s\n", names[m], cudaGetErrorString(e)); return 1; }
      printf("%-16s %2d warps/SM: %.2f warp-lanes/clk/SM (instr/clk/SM x32)\n", names[m], threads / 32, (double)iters * 16 * threads / cyc[0]);
    }
  }
  return 0;
}
