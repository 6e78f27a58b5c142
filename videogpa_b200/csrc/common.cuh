// Shared device/host helpers for the videogpa_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace vgpa {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define VGPA_CHECK(cond, ...)                       \
  do {                                              \
    if (!(cond)) {                                  \
      ::vgpa::set_error(__VA_ARGS__);               \
      return 1;                                     \
    }                                               \
  } while (0)

#define VGPA_CUDA(expr)                                             \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return ::vgpa::cuda_fail(_e, #expr);     \
  } while (0)

#define VGPA_LAUNCH_CHECK(name)                                     \
  do {                                                              \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return ::vgpa::cuda_fail(_e, name);      \
  } while (0)

int num_sms();
// vae_norm.cu: fold [nblocks][2][C] per-block channel sums / sums of squares into mean_rstd [2][groups] (fp64, fixed order)
int launch_gn_finalize(const float* partial, int nblocks, int C, int groups, double count, float eps, float* mean_rstd, cudaStream_t stream);

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace vgpa
