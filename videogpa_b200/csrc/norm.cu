// K3: fused LayerNorm + adaLN modulation over the [text; video] token rows of the DiT.
//
// Replaces CogVideoXLayerNormZero.forward / AdaLayerNorm.forward / norm_final of diffusers'
// CogVideoXTransformer3DModel (`LN(x) * (1 + scale)[:, None] + shift[:, None]`,
// one LayerNorm module shared by the text and the video rows, separate shift/scale per segment).
// HBM-bound: one read and one write of a [rows, D] bf16 tensor; one warp owns one row, 128-bit
// loads, statistics in fp32 (two-pass over registers), roundings placed where eager bf16 has them.
#include "common.cuh"
#include "../../include/videogpa_b200.h"

namespace vgpa {
namespace {

constexpr int LN_WARPS = 8;

struct LnParams {
  const void* x;        // bf16, or float when XF32
  __nv_bfloat16* out;
  long long ldx, ldo;
  int rows, D;
  const __nv_bfloat16* w;
  const __nv_bfloat16* b;
  float eps;
  int rows_per_sample, text_rows;
  const __nv_bfloat16* shift_txt;
  const __nv_bfloat16* scale_txt;
  const __nv_bfloat16* shift_vid;
  const __nv_bfloat16* scale_vid;
  long long mod_stride_b;
};

// VPL = uint4 vectors per lane (D = VPL * 256)
// XF32: x is the fp32 residual stream of the Wan2.2 forward; everything up to the single bf16 rounding of the output is fp32
// (torch.autocast semantics). Otherwise x is bf16 and the modulation follows torch's eager-bf16 roundings.
// Two CTAs per SM up to D = 3072 (VPL 12): the bf16 row kernel needed 130 registers, two more than lets a second 256-thread CTA
// onto the SM, and the kernel is bound by the rows it keeps in flight (8 warps = 8 rows = 48 KB per SM is half of what HBM needs).
template <int VPL, bool XF32>
__global__ void __launch_bounds__(LN_WARPS * 32, (VPL <= 12 && !XF32) ? 2 : 1)
ln_modulate_kernel(LnParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + warp;
  if (row >= p.rows) return;
  float v[VPL * 8];
  if constexpr (XF32) {
    const float4* xr = reinterpret_cast<const float4*>(static_cast<const float*>(p.x) + static_cast<long long>(row) * p.ldx);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 a = xr[(i * 32 + lane) * 2], b = xr[(i * 32 + lane) * 2 + 1];
      v[i * 8 + 0] = a.x; v[i * 8 + 1] = a.y; v[i * 8 + 2] = a.z; v[i * 8 + 3] = a.w;
      v[i * 8 + 4] = b.x; v[i * 8 + 5] = b.y; v[i * 8 + 6] = b.z; v[i * 8 + 7] = b.w;
    }
  } else {
    const uint4* xr = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.x) + static_cast<long long>(row) * p.ldx);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const uint4 u = xr[i * 32 + lane];
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      v[i * 8 + 0] = a.x; v[i * 8 + 1] = a.y; v[i * 8 + 2] = b.x; v[i * 8 + 3] = b.y;
      v[i * 8 + 4] = c.x; v[i * 8 + 5] = c.y; v[i * 8 + 6] = d.x; v[i * 8 + 7] = d.y;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) s += v[i];
  const float mean = warp_sum(s) / static_cast<float>(p.D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(p.D) + p.eps);

  int b = 0, srow = row;
  if (p.rows_per_sample > 0) { b = row / p.rows_per_sample; srow = row - b * p.rows_per_sample; }
  const bool is_txt = srow < p.text_rows;
  const __nv_bfloat16* shift = is_txt ? p.shift_txt : p.shift_vid;
  const __nv_bfloat16* scale = is_txt ? p.scale_txt : p.scale_vid;
  const uint4* sh4 = shift ? reinterpret_cast<const uint4*>(shift + b * p.mod_stride_b) : nullptr;
  const uint4* sc4 = scale ? reinterpret_cast<const uint4*>(scale + b * p.mod_stride_b) : nullptr;
  const uint4* w4 = p.w ? reinterpret_cast<const uint4*>(p.w) : nullptr;
  const uint4* b4 = p.b ? reinterpret_cast<const uint4*>(p.b) : nullptr;
  uint4* orow = reinterpret_cast<uint4*>(p.out + static_cast<long long>(row) * p.ldo);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int idx = i * 32 + lane;
    float y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) y[k] = (v[i * 8 + k] - mean) * rstd;
    if (w4) {
      const uint4 wu = __ldg(w4 + idx);
      const uint32_t wv[4] = {wu.x, wu.y, wu.z, wu.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(wv[k]); y[2 * k] *= f.x; y[2 * k + 1] *= f.y; }
    }
    if (b4) {
      const uint4 bu = __ldg(b4 + idx);
      const uint32_t bv[4] = {bu.x, bu.y, bu.z, bu.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(bv[k]); y[2 * k] += f.x; y[2 * k + 1] += f.y; }
    }
    if constexpr (XF32) {
      if (sc4) {
        const uint4 su = __ldg(sc4 + idx);
        const uint32_t sv[4] = {su.x, su.y, su.z, su.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(sv[k]); y[2 * k] *= 1.0f + f.x; y[2 * k + 1] *= 1.0f + f.y; }
      }
      if (sh4) {
        const uint4 su = __ldg(sh4 + idx);
        const uint32_t sv[4] = {su.x, su.y, su.z, su.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16x2(sv[k]); y[2 * k] += f.x; y[2 * k + 1] += f.y; }
      }
      uint4 o;
      o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]); o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
      orow[idx] = o;
      continue;
    }
    // eager bf16 semantics with packed bf16x2 hardware ops (one rounding per op, exactly what torch's bf16 kernels do):
    // n = bf16(LN(x));  n = n * bf16(1 + scale);  n = n + shift
    __nv_bfloat162 nb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) nb[k] = __floats2bfloat162_rn(y[2 * k], y[2 * k + 1]);
    if (sc4) {
      const uint4 su = __ldg(sc4 + idx);
      const uint32_t sv[4] = {su.x, su.y, su.z, su.w};
      const __nv_bfloat162 one = __floats2bfloat162_rn(1.0f, 1.0f);
#pragma unroll
      for (int k = 0; k < 4; ++k) nb[k] = __hmul2(nb[k], __hadd2(one, *reinterpret_cast<const __nv_bfloat162*>(&sv[k])));
    }
    if (sh4) {
      const uint4 su = __ldg(sh4 + idx);
      const uint32_t sv[4] = {su.x, su.y, su.z, su.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) nb[k] = __hadd2(nb[k], *reinterpret_cast<const __nv_bfloat162*>(&sv[k]));
    }
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&nb[0]); o.y = *reinterpret_cast<uint32_t*>(&nb[1]);
    o.z = *reinterpret_cast<uint32_t*>(&nb[2]); o.w = *reinterpret_cast<uint32_t*>(&nb[3]);
    orow[idx] = o;
  }
}

}  // namespace
}  // namespace vgpa

extern "C" int vgpa_layernorm_modulate_bf16(const vgpa_layernorm_args* a, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr, "vgpa_layernorm_modulate_bf16: null args");
  VGPA_CHECK(a->rows > 0 && a->D > 0, "vgpa_layernorm_modulate_bf16: bad shape rows=%d D=%d", a->rows, a->D);
  VGPA_CHECK(a->D % 256 == 0 && a->D <= 4096, "vgpa_layernorm_modulate_bf16: D=%d must be a multiple of 256, <= 4096", a->D);
  VGPA_CHECK(a->x && a->out, "vgpa_layernorm_modulate_bf16: null tensor pointer");
  VGPA_CHECK(a->ldx % 8 == 0 && a->ldo % 8 == 0 && a->mod_stride_b % 8 == 0, "vgpa_layernorm_modulate_bf16: strides must be multiples of 8");
  VGPA_CHECK((a->shift_txt == nullptr) == (a->shift_vid == nullptr) && (a->scale_txt == nullptr) == (a->scale_vid == nullptr),
             "vgpa_layernorm_modulate_bf16: text/video modulation pointers must be set together");
  LnParams p;
  p.x = a->x;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.ldx = a->ldx; p.ldo = a->ldo; p.rows = a->rows; p.D = a->D;
  p.w = static_cast<const __nv_bfloat16*>(a->ln_weight);
  p.b = static_cast<const __nv_bfloat16*>(a->ln_bias);
  p.eps = a->eps;
  p.rows_per_sample = a->rows_per_sample; p.text_rows = a->text_rows;
  p.shift_txt = static_cast<const __nv_bfloat16*>(a->shift_txt);
  p.scale_txt = static_cast<const __nv_bfloat16*>(a->scale_txt);
  p.shift_vid = static_cast<const __nv_bfloat16*>(a->shift_vid);
  p.scale_vid = static_cast<const __nv_bfloat16*>(a->scale_vid);
  p.mod_stride_b = a->mod_stride_b;
  const int grid = (a->rows + LN_WARPS - 1) / LN_WARPS;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a->D / 256) {
#define VGPA_LN_CASE(V)                                                              \
  case V:                                                                          \
    if (a->x_is_f32) ln_modulate_kernel<V, true><<<grid, LN_WARPS * 32, 0, s>>>(p);  \
    else ln_modulate_kernel<V, false><<<grid, LN_WARPS * 32, 0, s>>>(p);             \
    break;
    VGPA_LN_CASE(1) VGPA_LN_CASE(2) VGPA_LN_CASE(3) VGPA_LN_CASE(4) VGPA_LN_CASE(5) VGPA_LN_CASE(6)
    VGPA_LN_CASE(7) VGPA_LN_CASE(8) VGPA_LN_CASE(9) VGPA_LN_CASE(10) VGPA_LN_CASE(11) VGPA_LN_CASE(12)
    VGPA_LN_CASE(13) VGPA_LN_CASE(14) VGPA_LN_CASE(15) VGPA_LN_CASE(16)
#undef VGPA_LN_CASE
    default:
      set_error("vgpa_layernorm_modulate_bf16: unsupported D=%d", a->D);
      return 1;
  }
  VGPA_LAUNCH_CHECK("ln_modulate_kernel");
  return 0;
}
