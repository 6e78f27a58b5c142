// Small HBM-/latency-bound pieces around the DiT block stack:
//   - vgpa_linear_smallm_bf16 : out[M<=8, N] = bias + act_in(x) W^T. The conditioning path — timestep MLP
//     (time_embedding.linear_1/2) and every adaLN `linear(SiLU(emb))` — has M = batch, so it is a
//     weight-streaming GEMV: one warp per output feature, 128-bit weight loads, fp32 accumulate.
//   - vgpa_timestep_embedding_bf16 : sinusoid(timestep) with flip_sin_to_cos (cos first), fp32 math.
//   - vgpa_patchify_bf16 / vgpa_unpatchify_bf16 : the 2x2 patch gather feeding patch_embed.proj as a
//     K = C*4 GEMM, and the inverse scatter after proj_out.
#include "common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {

constexpr int SM_MAX_M = 8;

__device__ __forceinline__ float silu_bf16(float x) {
  // eager bf16 SiLU: fp32 math on the bf16 input, one rounding on the way out
  return bf16_round(x / (1.0f + expf(-x)));
}

template <int ACT>
__global__ void __launch_bounds__(256)
linear_smallm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ W,
                     const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, int M,
                     int N, int K, long long ldx, long long ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  float acc[SM_MAX_M];
#pragma unroll
  for (int m = 0; m < SM_MAX_M; ++m) acc[m] = 0.f;
  const uint4* wrow = reinterpret_cast<const uint4*>(W + static_cast<long long>(n) * K);
  for (int kv = lane; kv < K / 8; kv += 32) {
    const uint4 wu = __ldg(wrow + kv);
    const uint32_t wv[4] = {wu.x, wu.y, wu.z, wu.w};
    float wf[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = unpack_bf16x2(wv[i]); wf[2 * i] = f.x; wf[2 * i + 1] = f.y; }
#pragma unroll
    for (int m = 0; m < SM_MAX_M; ++m) {
      if (m < M) {
        const uint4 xu = *reinterpret_cast<const uint4*>(x + m * ldx + kv * 8);
        const uint32_t xv[4] = {xu.x, xu.y, xu.z, xu.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 f = unpack_bf16x2(xv[i]);
          if (ACT == 1) { f.x = silu_bf16(f.x); f.y = silu_bf16(f.y); }
          acc[m] = fmaf(f.x, wf[2 * i], acc[m]);
          acc[m] = fmaf(f.y, wf[2 * i + 1], acc[m]);
        }
      }
    }
  }
  const float bv = bias ? __bfloat162float(bias[n]) : 0.f;
#pragma unroll
  for (int m = 0; m < SM_MAX_M; ++m) {
    if (m < M) {
      const float s = warp_sum(acc[m]);
      if (lane == 0) out[m * ldo + n] = __float2bfloat16_rn(s + bv);
    }
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out,
                                          int B, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i - b * half;
  const float freq = expf(-9.210340371976184f * static_cast<float>(k) / static_cast<float>(half));
  const float ang = t[b] * freq;
  out[static_cast<long long>(b) * dim + k] = __float2bfloat16_rn(cosf(ang));
  out[static_cast<long long>(b) * dim + half + k] = __float2bfloat16_rn(sinf(ang));
}

// in [BF, C, H, W] -> out [BF * (H/2) * (W/2), C*4], feature = c*4 + ph*2 + pw  (Conv2d k=2 s=2 weight order)
__global__ void patchify_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                long long total, int C, int H, int W) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int feat = C * 4;
  const int f = static_cast<int>(i % feat);
  const long long tok = i / feat;
  const int hw = (H / 2) * (W / 2);
  const int bf = static_cast<int>(tok / hw);
  const int ij = static_cast<int>(tok - static_cast<long long>(bf) * hw);
  const int ti = ij / (W / 2), tj = ij - ti * (W / 2);
  const int c = f >> 2, ph = (f >> 1) & 1, pw = f & 1;
  out[i] = in[((static_cast<long long>(bf) * C + c) * H + 2 * ti + ph) * W + 2 * tj + pw];
}

// in [BF * h * w, ldi] with feature = c*4 + ph*2 + pw -> out [BF, C, 2h, 2w]
__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                  long long total, int C, int H, int W, long long ldi) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % W);
  const int y = static_cast<int>((i / W) % H);
  const int c = static_cast<int>((i / (static_cast<long long>(W) * H)) % C);
  const long long bf = i / (static_cast<long long>(W) * H * C);
  const long long tok = (bf * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1);
  out[i] = in[tok * ldi + c * 4 + (y & 1) * 2 + (x & 1)];
}

}  // namespace
}  // namespace vgpa

extern "C" int vgpa_linear_smallm_bf16(const void* x, const void* W, const void* bias, void* out, int M, int N,
                                       int K, int64_t ldx, int64_t ldo, int act_in, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(x && W && out, "vgpa_linear_smallm_bf16: null tensor pointer");
  VGPA_CHECK(M >= 1 && M <= SM_MAX_M, "vgpa_linear_smallm_bf16: M=%d must be in [1, %d]", M, SM_MAX_M);
  VGPA_CHECK(N > 0 && K > 0 && K % 8 == 0 && ldx % 8 == 0, "vgpa_linear_smallm_bf16: bad N=%d K=%d ldx=%lld", N, K, (long long)ldx);
  VGPA_CHECK(act_in == VGPA_ACT_NONE || act_in == VGPA_ACT_SILU, "vgpa_linear_smallm_bf16: unknown act_in %d", act_in);
  const int grid = (N + 7) / 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (act_in == VGPA_ACT_SILU)
    linear_smallm_kernel<1><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(W),
                                                 static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(out), M, N, K, ldx, ldo);
  else
    linear_smallm_kernel<0><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(W),
                                                 static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(out), M, N, K, ldx, ldo);
  VGPA_LAUNCH_CHECK("linear_smallm_kernel");
  return 0;
}

extern "C" int vgpa_timestep_embedding_bf16(const float* d_timesteps, void* out, int B, int dim, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(d_timesteps && out, "vgpa_timestep_embedding_bf16: null pointer");
  VGPA_CHECK(B > 0 && dim > 0 && dim % 2 == 0, "vgpa_timestep_embedding_bf16: bad B=%d dim=%d", B, dim);
  const int n = B * (dim / 2);
  timestep_embedding_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_timesteps, static_cast<__nv_bfloat16*>(out), B, dim);
  VGPA_LAUNCH_CHECK("timestep_embedding_kernel");
  return 0;
}

extern "C" int vgpa_patchify_bf16(const void* in, void* out, int BF, int C, int H, int W, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(in && out, "vgpa_patchify_bf16: null pointer");
  VGPA_CHECK(BF > 0 && C > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "vgpa_patchify_bf16: bad shape BF=%d C=%d H=%d W=%d", BF, C, H, W);
  const long long total = static_cast<long long>(BF) * C * H * W;
  patchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), total, C, H, W);
  VGPA_LAUNCH_CHECK("patchify_kernel");
  return 0;
}

extern "C" int vgpa_unpatchify_bf16(const void* in, void* out, int BF, int C, int H, int W, int64_t ldi, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(in && out, "vgpa_unpatchify_bf16: null pointer");
  VGPA_CHECK(BF > 0 && C > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && ldi >= C * 4, "vgpa_unpatchify_bf16: bad shape");
  const long long total = static_cast<long long>(BF) * C * H * W;
  unpatchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), total, C, H, W, ldi);
  VGPA_LAUNCH_CHECK("unpatchify_kernel");
  return 0;
}
