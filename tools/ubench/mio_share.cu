// Microbenchmark: do tcgen05.ld (LDTM) and MUFU.EX2 contend for the same issue path (MIO) on an SM sub-partition?
// 8 warps per SM: warps 0-3 run a MUFU loop, warps 4-7 a TMEM-load loop (one of each per sub-partition). Each group is timed
// alone and together.
#include "sm100.cuh"
#include <cstdio>
using namespace vgpa;

__global__ void mix(int iters, int run_mufu, int run_ld, unsigned long long* cyc_mufu, unsigned long long* cyc_ld, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = 0.5f + k * 0.001f + threadIdx.x * 1e-6f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < 4) {
    if (run_mufu)
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
      }
  } else {
    if (run_ld)
      for (int i = 0; i < iters / 8; ++i) {
        uint32_t r[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) ptx::tmem_ld_32x32(base + c * 32, r[c]);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int k = 0; k < 32; ++k) acc ^= r[c][k];
      }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) {
    if (warp == 0) cyc_mufu[blockIdx.x] = t1 - t0;
    if (warp == 4) cyc_ld[blockIdx.x] = t1 - t0;
  }
  float f = 0; for (int k = 0; k < 16; ++k) f += x[k];
  if (f == 1.2345f || acc == 0x12345u) sink[0] = f;
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(slot, 512);
}

int main() {
  unsigned long long *cm, *cl; float* sink;
  cudaMallocManaged(&cm, 148 * 8); cudaMallocManaged(&cl, 148 * 8); cudaMalloc(&sink, 64);
  const int iters = 4000;
  for (int mode = 0; mode < 3; ++mode) {
    const int rm = mode != 1, rl = mode != 0;
    mix<<<148, 256>>>(iters, rm, rl, cm, cl, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return 1; }
    const double mufu_ops = (double)iters * 16 * 128;            // per SM (4 warps x 32 lanes)
    const double ld_bytes = (double)(iters / 8) * 4 * 4096 * 4;  // per SM (4 warps)
    printf("%s:", mode == 0 ? "MUFU only      " : mode == 1 ? "LDTM only      " : "MUFU + LDTM    ");
    if (rm) printf("  MUFU %.2f ops/clk/SM (%llu cycles)", mufu_ops / cm[0], cm[0]);
    if (rl) printf("  LDTM %.1f B/clk/SM (%llu cycles, %.0f cycles per x32 load per warp)", ld_bytes / cl[0], cl[0], (double)cl[0] / (iters / 8) / 4);
    printf("\n");
  }
  return 0;
}
