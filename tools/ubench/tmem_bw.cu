// Microbenchmark: TMEM -> register bandwidth of tcgen05.ld (32x32b.x32) per SM, and MUFU.EX2 / FFMA2 rates.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I videogpa_b200/csrc -I include tools/ubench/tmem_bw.cu -o /tmp/tmem_bw
#include "sm100.cuh"
#include <cstdio>
using namespace vgpa;

__global__ void tmem_ld_bw(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c) ptx::tmem_ld_32x32(base + c * 32, r[c]);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < 32; ++k) acc ^= r[c][k];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345) sink[0] = acc;
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(slot, 512);
}

__global__ void mufu_rate(int iters, unsigned long long* cycles, float* sink, float seed) {
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = seed + k * 0.001f + threadIdx.x * 1e-6f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = ptx::ex2_approx(x[k]);
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0; for (int k = 0; k < 16; ++k) s += x[k];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (s == 1.2345f) sink[0] = s;
}

__global__ void ffma2_rate(int iters, unsigned long long* cycles, float* sink, float seed) {
  uint64_t x[16];
  uint64_t a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(seed), "f"(seed * 0.5f));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(seed * 0.25f), "f"(seed * 0.125f));
#pragma unroll
  for (int k = 0; k < 16; ++k) asm("mov.b64 %0, {%1, %2};" : "=l"(x[k]) : "f"(seed + k), "f"(seed - k));
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(a), "l"(b));
  }
  __syncthreads();
  const long long t1 = clock64();
  uint64_t s = 0; for (int k = 0; k < 16; ++k) s ^= x[k];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (s == 12345ull) sink[0] = 1.f;
}

__global__ void fmnmx3_rate(int iters, unsigned long long* cycles, float* sink, float seed) {
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = seed + k;
  float a = seed * 0.5f, b = seed * 0.25f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[k]) : "f"(a), "f"(b));
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0; for (int k = 0; k < 16; ++k) s += x[k];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (s == 1.2345f) sink[0] = s;
}

int main() {
  unsigned long long* cyc; uint32_t* sink;
  cudaMallocManaged(&cyc, 148 * 8); cudaMalloc(&sink, 64);
  const int iters = 2000;
  for (int threads : {32, 64, 128, 256}) {
    tmem_ld_bw<<<148, threads>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tmem_ld_bw failed: %s\n", cudaGetErrorString(e)); return 1; }
    double bytes = (double)iters * (threads / 32) * 4 * 4096;
    printf("tmem_ld 32x32b.x32: %d warps/SM: %.1f cycles/iter, %.1f B/clk/SM\n", threads / 32, (double)cyc[0] / iters, bytes / cyc[0]);
  }
  for (int threads : {128, 256, 512}) {
    mufu_rate<<<148, threads>>>(iters, cyc, (float*)sink, 0.5f); cudaDeviceSynchronize();
    printf("MUFU.EX2: %d warps/SM: %.2f ops/clk/SM\n", threads / 32, (double)iters * 16 * threads / cyc[0]);
    ffma2_rate<<<148, threads>>>(iters, cyc, (float*)sink, 0.5f); cudaDeviceSynchronize();
    printf("FFMA2: %d warps/SM: %.2f packed-instr lanes/clk/SM (x2 flops-pairs)\n", threads / 32, (double)iters * 16 * threads / cyc[0]);
    fmnmx3_rate<<<148, threads>>>(iters, cyc, (float*)sink, 0.5f); cudaDeviceSynchronize();
    printf("FMNMX3: %d warps/SM: %.2f lanes/clk/SM\n", threads / 32, (double)iters * 16 * threads / cyc[0]);
  }
  return 0;
}
