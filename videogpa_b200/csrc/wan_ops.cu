// Row-wise helpers of the Wan2.2 DiT block (reference call sites
// generate/Wan2.2-TI2V-5B.py:120-129, train/Wan2.2-TI2V-5B/03_train.py:228-233):
//   * WanRMSNorm over the full model dim on q / k (qk_norm) fused with the complex-pair 3-D RoPE of
//     WanSelfAttention (rope_apply), in place on the fused projection buffer;
//   * modulation + time-projection sum of WanAttentionBlock / Head.
// HBM-bound: one warp owns one row, 128-bit loads, fp32 statistics.
#include "common.cuh"
#include "../../include/videogpa_b200.h"

namespace vgpa {
namespace {

constexpr int RN_WARPS = 8;

struct RnParams {
  __nv_bfloat16* x;
  long long ldx;
  int rows, D, head_dim, rows_per_sample;
  const float* weight;
  float eps;
  const float* cos;
  const float* sin;
};

// VPL = uint4 vectors per lane (D = VPL * 256)
template <int VPL>
__global__ void __launch_bounds__(RN_WARPS * 32, (VPL <= 12) ? 2 : 1)   // 144 registers at VPL 12 kept a second CTA off the SM (cf. ln_modulate_kernel)
rmsnorm_rope_kernel(RnParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * RN_WARPS + warp;
  if (row >= p.rows) return;
  uint4* xr = reinterpret_cast<uint4*>(p.x + static_cast<long long>(row) * p.ldx);
  float v[VPL * 8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const uint4 u = xr[i * 32 + lane];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(w[k]);
      v[i * 8 + 2 * k] = f.x; v[i * 8 + 2 * k + 1] = f.y;
      ss += f.x * f.x + f.y * f.y;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(p.D) + p.eps);
  const int srow = p.rows_per_sample > 0 ? row % p.rows_per_sample : row;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c0 = (i * 32 + lane) * 8;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.weight + c0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.weight + c0 + 4));
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    float y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) y[k] = bf16_round(v[i * 8 + k] * rstd) * wv[k];   // `_norm(x.float()).type_as(x) * weight`
    if (p.cos != nullptr) {
      const int d0 = c0 % p.head_dim;                                             // 8 consecutive dims of one head
      const float* cs = p.cos + static_cast<long long>(srow) * p.head_dim + d0;
      const float* sn = p.sin + static_cast<long long>(srow) * p.head_dim + d0;
      const float4 ca = __ldg(reinterpret_cast<const float4*>(cs)), cb = __ldg(reinterpret_cast<const float4*>(cs + 4));
      const float4 sa = __ldg(reinterpret_cast<const float4*>(sn)), sb = __ldg(reinterpret_cast<const float4*>(sn + 4));
      const float cv[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
      const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
      float r[8];
#pragma unroll
      for (int k = 0; k < 8; k += 2) {                                            // (a + ib)(cos + i sin)
        r[k] = y[k] * cv[k] - y[k + 1] * sv[k];
        r[k + 1] = y[k + 1] * cv[k + 1] + y[k] * sv[k + 1];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) y[k] = r[k];
    }
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
    xr[i * 32 + lane] = o;
  }
}

__global__ void add_rows_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ b, __nv_bfloat16* __restrict__ out,
                                int R, long long N, long long lda) {
  const long long n = static_cast<long long>(R) * N;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / N, c = i - r * N;
    out[i] = __float2bfloat16_rn(__bfloat162float(a[r * lda + c]) + b[c]);
  }
}

}  // namespace
}  // namespace vgpa

extern "C" int vgpa_rmsnorm_rope_bf16(void* x, int rows, int D, int64_t ldx, const float* weight, float eps,
                                      const float* rope_cos, const float* rope_sin, int head_dim, int rows_per_sample,
                                      void* stream) {
  using namespace vgpa;
  VGPA_CHECK(x && weight, "vgpa_rmsnorm_rope_bf16: null pointer");
  VGPA_CHECK(rows > 0 && D > 0 && D % 256 == 0 && D <= 4096, "vgpa_rmsnorm_rope_bf16: D=%d must be a multiple of 256, <= 4096", D);
  VGPA_CHECK(ldx % 8 == 0 && ldx >= D, "vgpa_rmsnorm_rope_bf16: ldx invalid");
  VGPA_CHECK((rope_cos == nullptr) == (rope_sin == nullptr), "vgpa_rmsnorm_rope_bf16: cos/sin must both be set or both null");
  VGPA_CHECK(rope_cos == nullptr || (head_dim > 0 && head_dim % 8 == 0 && D % head_dim == 0), "vgpa_rmsnorm_rope_bf16: bad head_dim %d", head_dim);
  RnParams p;
  p.x = static_cast<__nv_bfloat16*>(x);
  p.ldx = ldx; p.rows = rows; p.D = D; p.head_dim = head_dim > 0 ? head_dim : D; p.rows_per_sample = rows_per_sample;
  p.weight = weight; p.eps = eps; p.cos = rope_cos; p.sin = rope_sin;
  const int grid = (rows + RN_WARPS - 1) / RN_WARPS;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (D / 256) {
#define VGPA_RN_CASE(V) case V: rmsnorm_rope_kernel<V><<<grid, RN_WARPS * 32, 0, s>>>(p); break;
    VGPA_RN_CASE(1) VGPA_RN_CASE(2) VGPA_RN_CASE(3) VGPA_RN_CASE(4) VGPA_RN_CASE(5) VGPA_RN_CASE(6)
    VGPA_RN_CASE(7) VGPA_RN_CASE(8) VGPA_RN_CASE(9) VGPA_RN_CASE(10) VGPA_RN_CASE(11) VGPA_RN_CASE(12)
    VGPA_RN_CASE(13) VGPA_RN_CASE(14) VGPA_RN_CASE(15) VGPA_RN_CASE(16)
#undef VGPA_RN_CASE
    default:
      set_error("vgpa_rmsnorm_rope_bf16: unsupported D=%d", D);
      return 1;
  }
  VGPA_LAUNCH_CHECK("rmsnorm_rope_kernel");
  return 0;
}

extern "C" int vgpa_add_rows_bf16(const void* a, const float* b, void* out, int R, int64_t N, int64_t lda, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a && b && out && R > 0 && N > 0 && lda >= N, "vgpa_add_rows_bf16: bad arguments");
  const long long n = static_cast<long long>(R) * N;
  int grid = static_cast<int>((n + 255) / 256);
  if (grid > 1184) grid = 1184;
  add_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(a), b,
                                                                         static_cast<__nv_bfloat16*>(out), R, N, lda);
  VGPA_LAUNCH_CHECK("add_rows_kernel");
  return 0;
}
