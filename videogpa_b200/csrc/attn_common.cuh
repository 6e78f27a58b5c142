// Softmax arithmetic shared by the attention kernels: packed fp32x2 FMA-pipe ops, 3-input max and the
// Cody-Waite + degree-3 polynomial exp2 that offloads part of the exponentials from MUFU to the FMA pipe.
#pragma once
#include "sm100.cuh"

namespace vgpa {
namespace attn {

// ------------------------------------------------------------------ packed fp32x2 helpers (FFMA2 / FADD2)
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_add_rm(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// 2^x for a packed pair on the FMA pipe: floor via round-to-minus-inf magic add, degree-3 minimax
// polynomial of 2^f on [0,1), exponent spliced in with integer shift+add. x <= 127 is guaranteed by
// the caller (x <= AT_RESCALE_THRESHOLD); x is clamped at -127 from below.
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  f2_unpack(x2, x0, x1);
  x0 = fmaxf(x0, -127.0f);
  x1 = fmaxf(x1, -127.0f);
  const uint64_t xc = f2_pack(x0, x1);
  const uint64_t magic = f2_pack(12582912.0f, 12582912.0f);                 // 2^23 + 2^22
  const uint64_t xr = f2_add_rm(xc, magic);                                  // low mantissa bits = floor(x)
  const uint64_t fl = f2_sub(xr, magic);
  const uint64_t fr = f2_sub(xc, fl);                                        // in [0, 1)
  uint64_t acc = f2_fma(fr, f2_pack(0.077119089663028717f, 0.077119089663028717f),
                        f2_pack(0.227564394474029541f, 0.227564394474029541f));
  acc = f2_fma(acc, fr, f2_pack(0.695146143436431885f, 0.695146143436431885f));
  acc = f2_fma(acc, fr, f2_pack(1.0f, 1.0f));
  float r0, r1, q0, q1;
  f2_unpack(xr, r0, r1);
  f2_unpack(acc, q0, q1);
  p0 = __int_as_float((__float_as_int(r0) << 23) + __float_as_int(q0));
  p1 = __int_as_float((__float_as_int(r1) << 23) + __float_as_int(q1));
}

// Bounded-softmax dispatch (attention_d64b_sm100.cu): M = ceil(max|q| max|k| |scale log2e| + 0.5) of a (batch, head) from the
// pre-pass buffer {max|q|^2, max|k|^2}. Heads with M <= kBoundedMax are served by the bounded kernel, the rest by the
// exact online-softmax kernel; both kernels evaluate this same expression.
constexpr float kBoundedMax = 90.0f;
__device__ __forceinline__ float bounded_m(const float* bounds, int bh, float scale_log2) {
  return ceilf(sqrtf(bounds[2 * bh]) * sqrtf(bounds[2 * bh + 1]) * fabsf(scale_log2) + 0.5f);
}

// NPOLY of the 64 column pairs of a row take the polynomial path, spread evenly.
template <int NPOLY>
__host__ __device__ constexpr bool pair_uses_poly(int pi) {
  return ((pi + 1) * NPOLY) / 64 != (pi * NPOLY) / 64;
}

}  // namespace attn
}  // namespace vgpa
