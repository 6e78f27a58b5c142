// Row-wise / elementwise backward kernels of the DiT block for the DPO training step (SURVEY.md §8 row f-2;
// train/CogVideoX-5B/03_train.py:116-157 back-propagates through CogVideoXBlock into the LoRA factors):
//   * backward of LayerNorm + adaLN modulation w.r.t. its input (the modulation vectors come from the frozen
//     conditioning path, so they receive no gradient), optionally added to the residual-stream gradient;
//   * per-head LayerNorm(64) on q / k (CogVideoXAttnProcessor2_0's norm_q / norm_k): forward out of place (the training
//     path keeps the raw projection for the backward) and backward;
//   * GELU(tanh) forward / backward on the stored pre-activation;
//   * multiplication of a row block by the per-sample, per-segment gate vector (backward of the gated residual).
// All are HBM-bound: 128-bit loads, one warp per row (or 8 lanes per head), fp32 statistics.
#include "common.cuh"
#include "../../include/videogpa_b200.h"

namespace vgpa {
namespace {

constexpr int TR_WARPS = 8;

struct LnBwdParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dy;
  const __nv_bfloat16* add;     // optional, added to the result
  __nv_bfloat16* dx;
  long long ldx, ld_dy, ld_add, ld_dx;
  int rows, D;
  const __nv_bfloat16* w;       // LN weight or null
  float eps;
  int rows_per_sample, text_rows;
  const __nv_bfloat16* scale_txt;
  const __nv_bfloat16* scale_vid;
  long long mod_stride_b;
};

template <int VPL>
__global__ void __launch_bounds__(TR_WARPS * 32)
ln_modulate_bwd_kernel(LnBwdParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * TR_WARPS + warp;
  if (row >= p.rows) return;
  const uint4* xr = reinterpret_cast<const uint4*>(p.x + static_cast<long long>(row) * p.ldx);
  const uint4* gr = reinterpret_cast<const uint4*>(p.dy + static_cast<long long>(row) * p.ld_dy);
  float v[VPL * 8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const uint4 u = xr[i * 32 + lane];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(w[k]); v[i * 8 + 2 * k] = f.x; v[i * 8 + 2 * k + 1] = f.y; }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) s += v[i];
  const float mean = warp_sum(s) / static_cast<float>(p.D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(p.D) + p.eps);
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) v[i] = (v[i] - mean) * rstd;                 // x_hat

  int b = 0, srow = row;
  if (p.rows_per_sample > 0) { b = row / p.rows_per_sample; srow = row - b * p.rows_per_sample; }
  const __nv_bfloat16* scale = (srow < p.text_rows) ? p.scale_txt : p.scale_vid;
  const uint4* sc4 = scale ? reinterpret_cast<const uint4*>(scale + b * p.mod_stride_b) : nullptr;
  const uint4* w4 = p.w ? reinterpret_cast<const uint4*>(p.w) : nullptr;

  // g = dy * (1 + scale) * w, kept as packed bf16x2 words? no: recomputed in the second pass from dy (L1-resident row)
  auto load_g = [&](int i, float (&g)[8]) {
    const int idx = i * 32 + lane;
    const uint4 du = gr[idx];
    const uint32_t dw[4] = {du.x, du.y, du.z, du.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(dw[k]); g[2 * k] = f.x; g[2 * k + 1] = f.y; }
    if (sc4) {
      const uint4 su = __ldg(sc4 + idx);
      const uint32_t sw[4] = {su.x, su.y, su.z, su.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(sw[k]); g[2 * k] *= 1.0f + f.x; g[2 * k + 1] *= 1.0f + f.y; }
    }
    if (w4) {
      const uint4 wu = __ldg(w4 + idx);
      const uint32_t ww[4] = {wu.x, wu.y, wu.z, wu.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(ww[k]); g[2 * k] *= f.x; g[2 * k + 1] *= f.y; }
    }
  };
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float g[8];
    load_g(i, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1 += g[k]; s2 = fmaf(g[k], v[i * 8 + k], s2); }
  }
  const float m1 = warp_sum(s1) / static_cast<float>(p.D), m2 = warp_sum(s2) / static_cast<float>(p.D);
  const uint4* ar = p.add ? reinterpret_cast<const uint4*>(p.add + static_cast<long long>(row) * p.ld_add) : nullptr;
  uint4* orow = reinterpret_cast<uint4*>(p.dx + static_cast<long long>(row) * p.ld_dx);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int idx = i * 32 + lane;
    float g[8], r[8];
    load_g(i, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = rstd * (g[k] - m1 - v[i * 8 + k] * m2);
    if (ar) {
      const uint4 au = ar[idx];
      const uint32_t aw[4] = {au.x, au.y, au.z, au.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(aw[k]); r[2 * k] += f.x; r[2 * k + 1] += f.y; }
    }
    uint4 o;
    o.x = pack_bf16x2(r[0], r[1]); o.y = pack_bf16x2(r[2], r[3]); o.z = pack_bf16x2(r[4], r[5]); o.w = pack_bf16x2(r[6], r[7]);
    orow[idx] = o;
  }
}

// ---------------------------------------------------------------------------------------------- per-head LayerNorm(64)
// 8 lanes per (row, head): one 16-byte vector each.
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

template <bool BWD>
__global__ void __launch_bounds__(256)
head_ln_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ out,
               long long n_groups, int heads, long long ldx, long long ld_dy, long long ldo, const float* __restrict__ w,
               const float* __restrict__ bias, float eps) {
  const long long gid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const bool live = gid < n_groups;
  const long long g = live ? gid : 0;
  const long long row = g / heads;
  const int h = static_cast<int>(g - row * heads);
  const long long off = h * 64 + sub * 8;
  const uint4 u = *reinterpret_cast<const uint4*>(x + row * ldx + off);
  const uint32_t xw[4] = {u.x, u.y, u.z, u.w};
  float v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(xw[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += v[k];
  const float mean = group8_sum(s) * (1.0f / 64.0f);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { const float d = v[k] - mean; q += d * d; }
  const float rstd = rsqrtf(group8_sum(q) * (1.0f / 64.0f) + eps);
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + sub * 8)), w1 = __ldg(reinterpret_cast<const float4*>(w + sub * 8 + 4));
  const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  float r[8];
  if (!BWD) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + sub * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + sub * 8 + 4));
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = (v[k] - mean) * rstd * wv[k] + bv[k];
  } else {
    const uint4 du = *reinterpret_cast<const uint4*>(dy + row * ld_dy + off);
    const uint32_t dw[4] = {du.x, du.y, du.z, du.w};
    float gg[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(dw[k]); gg[2 * k] = f.x * wv[2 * k]; gg[2 * k + 1] = f.y * wv[2 * k + 1]; }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = (v[k] - mean) * rstd; s1 += gg[k]; s2 = fmaf(gg[k], v[k], s2); }
    const float m1 = group8_sum(s1) * (1.0f / 64.0f), m2 = group8_sum(s2) * (1.0f / 64.0f);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = rstd * (gg[k] - m1 - v[k] * m2);
  }
  if (live) {
    uint4 o;
    o.x = pack_bf16x2(r[0], r[1]); o.y = pack_bf16x2(r[2], r[3]); o.z = pack_bf16x2(r[4], r[5]); o.w = pack_bf16x2(r[6], r[7]);
    *reinterpret_cast<uint4*>(out + row * ldo + off) = o;
  }
}

// ---------------------------------------------------------------------------------------------- GELU(tanh)
template <bool BWD>
__global__ void __launch_bounds__(256)
gelu_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ out, long long nvec) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 u = reinterpret_cast<const uint4*>(x)[i];
    const uint32_t xw[4] = {u.x, u.y, u.z, u.w};
    uint4 du = make_uint4(0, 0, 0, 0);
    if (BWD) du = reinterpret_cast<const uint4*>(dy)[i];
    const uint32_t dw[4] = {du.x, du.y, du.z, du.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 xf = unpack_bf16x2(xw[k]), df = unpack_bf16x2(dw[k]);
      const float xs[2] = {xf.x, xf.y}, ds[2] = {df.x, df.y};
      float r[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float xv = xs[e];
        const float t = tanhf(0.7978845608028654f * (xv + 0.044715f * xv * xv * xv));
        if (!BWD) {
          r[e] = 0.5f * xv * (1.0f + t);
        } else {
          const float dt = (1.0f - t * t) * 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * xv * xv);
          r[e] = ds[e] * (0.5f * (1.0f + t) + 0.5f * xv * dt);
        }
      }
      ow[k] = pack_bf16x2(r[0], r[1]);
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

// ---------------------------------------------------------------------------------------------- gate * rows
__global__ void __launch_bounds__(256)
scale_cols_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ out,
                  int rows, int nvec, long long ldx, long long ld_add, long long ldo, int rows_per_sample, int text_rows, const __nv_bfloat16* __restrict__ g_txt,
                  const __nv_bfloat16* __restrict__ g_vid, long long stride_b) {
  const long long total = static_cast<long long>(rows) * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / nvec);
    const int c = static_cast<int>(i - static_cast<long long>(row) * nvec) * 8;
    int b = 0, srow = row;
    if (rows_per_sample > 0) { b = row / rows_per_sample; srow = row - b * rows_per_sample; }
    const __nv_bfloat16* g = (srow < text_rows ? g_txt : g_vid) + b * stride_b + c;
    const uint4 xu = *reinterpret_cast<const uint4*>(x + row * ldx + c);
    const uint4 gu = __ldg(reinterpret_cast<const uint4*>(g));
    const __nv_bfloat162* xv = reinterpret_cast<const __nv_bfloat162*>(&xu);
    const __nv_bfloat162* gv = reinterpret_cast<const __nv_bfloat162*>(&gu);
    uint4 ou;
    __nv_bfloat162* ov = reinterpret_cast<__nv_bfloat162*>(&ou);
#pragma unroll
    for (int k = 0; k < 4; ++k) ov[k] = __hmul2_rn(xv[k], gv[k]);
    if (add != nullptr) {                                  // hidden = hidden + gate * branch: bf16 product, then bf16 sum
      const uint4 au = *reinterpret_cast<const uint4*>(add + row * ld_add + c);
      const __nv_bfloat162* av = reinterpret_cast<const __nv_bfloat162*>(&au);
#pragma unroll
      for (int k = 0; k < 4; ++k) ov[k] = __hadd2_rn(av[k], ov[k]);
    }
    *reinterpret_cast<uint4*>(out + row * ldo + c) = ou;
  }
}

// ---------------------------------------------------------------------------------------------- bf16 transpose
// out[c, r] = x[r, c] for r < rows, 0 for rows <= r < ldo (the padded token columns of a weight-gradient GEMM operand).
// 64 x 64 tiles through shared memory, 4-byte accesses on both sides.
constexpr int TP = 64;
__global__ void __launch_bounds__(256)
transpose_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int rows, int cols, long long ldx, long long ldo) {
  __shared__ __nv_bfloat16 tile[TP][TP + 2];
  const int r0 = blockIdx.y * TP, c0 = blockIdx.x * TP;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;          // 8 warps
  for (int i = w; i < TP; i += 8) {                                   // row r0 + i: lanes read column pairs
    const int r = r0 + i, c = c0 + 2 * lane;
    uint32_t v = 0;
    if (r < rows && c + 1 < cols) v = *reinterpret_cast<const uint32_t*>(x + r * ldx + c);
    else if (r < rows && c < cols) v = *reinterpret_cast<const unsigned short*>(x + r * ldx + c);
    *reinterpret_cast<uint32_t*>(&tile[i][2 * lane]) = v;
  }
  __syncthreads();
  for (int i = w; i < TP; i += 8) {                                   // output row c0 + i: lanes write row pairs r0 + 2 lane
    const int c = c0 + i, r = r0 + 2 * lane;
    if (c >= cols || r >= ldo) continue;
    const unsigned short lo = *reinterpret_cast<const unsigned short*>(&tile[2 * lane][i]);
    const unsigned short hi = *reinterpret_cast<const unsigned short*>(&tile[2 * lane + 1][i]);
    if (r + 1 < ldo) *reinterpret_cast<uint32_t*>(out + c * ldo + r) = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
    else *reinterpret_cast<unsigned short*>(out + c * ldo + r) = lo;
  }
}

inline unsigned grid_1d(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = 148ll * 16;
  return static_cast<unsigned>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace
}  // namespace vgpa

extern "C" int vgpa_layernorm_modulate_bwd_bf16(const vgpa_layernorm_args* a, const void* dy, int64_t ld_dy, const void* add,
                                                int64_t ld_add, void* dx, int64_t ld_dx, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr && a->x != nullptr && dy != nullptr && dx != nullptr, "vgpa_layernorm_modulate_bwd_bf16: null pointer");
  VGPA_CHECK(a->rows > 0 && a->D > 0 && a->D % 256 == 0 && a->D <= 4096, "vgpa_layernorm_modulate_bwd_bf16: D=%d must be a multiple of 256, <= 4096", a->D);
  VGPA_CHECK(a->ldx % 8 == 0 && ld_dy % 8 == 0 && ld_dx % 8 == 0 && a->ldx >= a->D && ld_dy >= a->D && ld_dx >= a->D,
             "vgpa_layernorm_modulate_bwd_bf16: bad leading dimension");
  VGPA_CHECK(add == nullptr || (ld_add % 8 == 0 && ld_add >= a->D), "vgpa_layernorm_modulate_bwd_bf16: bad ld_add");
  VGPA_CHECK((a->scale_txt == nullptr) == (a->scale_vid == nullptr), "vgpa_layernorm_modulate_bwd_bf16: scale_txt/scale_vid must both be set or both null");
  LnBwdParams p;
  p.x = static_cast<const __nv_bfloat16*>(a->x);
  p.dy = static_cast<const __nv_bfloat16*>(dy);
  p.add = static_cast<const __nv_bfloat16*>(add);
  p.dx = static_cast<__nv_bfloat16*>(dx);
  p.ldx = a->ldx; p.ld_dy = ld_dy; p.ld_add = ld_add; p.ld_dx = ld_dx;
  p.rows = a->rows; p.D = a->D;
  p.w = static_cast<const __nv_bfloat16*>(a->ln_weight);
  p.eps = a->eps;
  p.rows_per_sample = a->rows_per_sample; p.text_rows = a->text_rows;
  p.scale_txt = static_cast<const __nv_bfloat16*>(a->scale_txt);
  p.scale_vid = static_cast<const __nv_bfloat16*>(a->scale_vid);
  p.mod_stride_b = a->mod_stride_b;
  const int grid = (a->rows + TR_WARPS - 1) / TR_WARPS;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a->D / 256) {
#define VGPA_LNB_CASE(V) case V: ln_modulate_bwd_kernel<V><<<grid, TR_WARPS * 32, 0, s>>>(p); break;
    VGPA_LNB_CASE(1) VGPA_LNB_CASE(2) VGPA_LNB_CASE(3) VGPA_LNB_CASE(4) VGPA_LNB_CASE(5) VGPA_LNB_CASE(6)
    VGPA_LNB_CASE(7) VGPA_LNB_CASE(8) VGPA_LNB_CASE(9) VGPA_LNB_CASE(10) VGPA_LNB_CASE(11) VGPA_LNB_CASE(12)
    VGPA_LNB_CASE(13) VGPA_LNB_CASE(14) VGPA_LNB_CASE(15) VGPA_LNB_CASE(16)
#undef VGPA_LNB_CASE
    default:
      set_error("vgpa_layernorm_modulate_bwd_bf16: unsupported D=%d", a->D);
      return 1;
  }
  VGPA_LAUNCH_CHECK("ln_modulate_bwd_kernel");
  return 0;
}

extern "C" int vgpa_head_layernorm_bf16(const void* x, const void* dy, void* out, int64_t rows, int heads, int64_t ldx,
                                        int64_t ld_dy, int64_t ldo, const float* weight, const float* bias, float eps,
                                        int backward, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(x && out && weight, "vgpa_head_layernorm_bf16: null pointer");
  VGPA_CHECK(backward ? dy != nullptr : bias != nullptr, "vgpa_head_layernorm_bf16: %s", backward ? "dy is null" : "bias is null");
  VGPA_CHECK(rows > 0 && heads > 0, "vgpa_head_layernorm_bf16: bad shape");
  VGPA_CHECK(ldx % 8 == 0 && ldo % 8 == 0 && ldx >= heads * 64 && ldo >= heads * 64 && (!backward || (ld_dy % 8 == 0 && ld_dy >= heads * 64)),
             "vgpa_head_layernorm_bf16: leading dimensions must be multiples of 8 covering heads * 64 columns");
  const long long n_groups = static_cast<long long>(rows) * heads;
  const long long blocks = (n_groups * 8 + 255) / 256;
  VGPA_CHECK(blocks < (1ll << 31), "vgpa_head_layernorm_bf16: problem too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (backward)
    head_ln_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy),
                                                                     static_cast<__nv_bfloat16*>(out), n_groups, heads, ldx, ld_dy, ldo, weight, bias, eps);
  else
    head_ln_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), nullptr,
                                                                      static_cast<__nv_bfloat16*>(out), n_groups, heads, ldx, 0, ldo, weight, bias, eps);
  VGPA_LAUNCH_CHECK("head_ln_kernel");
  return 0;
}

extern "C" int vgpa_gelu_tanh_bf16(const void* x, const void* dy, void* out, int64_t n, int backward, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(x && out && (!backward || dy), "vgpa_gelu_tanh_bf16: null pointer");
  VGPA_CHECK(n > 0 && n % 8 == 0, "vgpa_gelu_tanh_bf16: n must be a positive multiple of 8");
  VGPA_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0,
             "vgpa_gelu_tanh_bf16: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (backward)
    gelu_kernel<true><<<grid_1d(n / 8, 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy),
                                                          static_cast<__nv_bfloat16*>(out), n / 8);
  else
    gelu_kernel<false><<<grid_1d(n / 8, 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), nullptr, static_cast<__nv_bfloat16*>(out), n / 8);
  VGPA_LAUNCH_CHECK("gelu_kernel");
  return 0;
}

extern "C" int vgpa_scale_cols_bf16(const void* x, const void* add, void* out, int rows, int D, int64_t ldx, int64_t ld_add,
                                    int64_t ldo, int rows_per_sample, int text_rows, const void* gate_txt, const void* gate_vid,
                                    int64_t gate_stride_b, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(x && out && gate_txt && gate_vid, "vgpa_scale_cols_bf16: null pointer");
  VGPA_CHECK(rows > 0 && D > 0 && D % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && ldx >= D && ldo >= D && gate_stride_b % 8 == 0,
             "vgpa_scale_cols_bf16: bad shape");
  VGPA_CHECK(add == nullptr || (ld_add % 8 == 0 && ld_add >= D), "vgpa_scale_cols_bf16: bad ld_add");
  const long long total = static_cast<long long>(rows) * (D / 8);
  scale_cols_kernel<<<grid_1d(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(out), rows, D / 8, ldx,
      ld_add, ldo, rows_per_sample, text_rows,
      static_cast<const __nv_bfloat16*>(gate_txt), static_cast<const __nv_bfloat16*>(gate_vid), gate_stride_b);
  VGPA_LAUNCH_CHECK("scale_cols_kernel");
  return 0;
}

extern "C" int vgpa_transpose_bf16(const void* x, void* out, int rows, int cols, int64_t ldx, int64_t ldo, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(x && out, "vgpa_transpose_bf16: null pointer");
  VGPA_CHECK(rows > 0 && cols > 0 && ldx >= cols && ldo >= rows, "vgpa_transpose_bf16: bad shape rows=%d cols=%d ldx=%lld ldo=%lld", rows, cols,
             (long long)ldx, (long long)ldo);
  VGPA_CHECK(ldx % 2 == 0 && ldo % 2 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 3) == 0,
             "vgpa_transpose_bf16: leading dimensions must be even and pointers 4-byte aligned");
  const dim3 grid(static_cast<unsigned>((cols + TP - 1) / TP), static_cast<unsigned>((ldo + TP - 1) / TP));
  VGPA_CHECK(grid.y < 65536, "vgpa_transpose_bf16: too many rows");
  transpose_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out),
                                                                          rows, cols, ldx, ldo);
  VGPA_LAUNCH_CHECK("transpose_kernel");
  return 0;
}
