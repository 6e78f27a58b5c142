// K5: Multi-View Consistency Score, all (clip, pair) units of a batch in three launches.
//
// Replaces MVCSMetric.compute (metrics/mvcs.py:12-114 of the reference; called from
// pipelines/process_video.py:185-192). The reference walks the T-1 consecutive pairs with ~25
// torch kernels and two host syncs per pair; here
//   1. mvcs_prepare : per pair, K_i^-1 and E_j E_i^-1 in fp64 (closed form / Gauss-Jordan), rounded to fp32
//   2. mvcs_pairs   : per pixel, back-project, move to camera j, project, bilinear-sample depth_j
//                     (grid_sample semantics: bilinear, zero padding, align_corners=True), masked
//                     squared error; per-block partial sums (fp64) and counts, no atomics
//   3. mvcs_finalize: fixed-order reduction of the partials, per-pair MSE, exp(-mean) per clip
// HBM-bound: 4 B/pixel/pair streamed (depth_i, float4) + the L2-resident bilinear gather of depth_j.
//
// Bit-exact mask counts at HBM speed. The in-image mask (mvcs.py:99) must match the reference / numpy oracle pixel for pixel,
// which fixes the fp32 operation order of the whole projection chain (no FMA contraction: this file is compiled with
// --fmad=false, IEEE division). Evaluated for every pixel that chain costs ~280 instructions (0.15 of the HBM roofline).
// Only pixels whose projection lands within rounding distance of a mask boundary can decide differently under another
// evaluation order, so the work is tiered:
//   tier 1 (mvcs_pairs_kernel body, ~97 % of the pixels): merged matrices per pair (M = K_j R K_i^-1 in fp64, rounded once),
//          h = d (M [u,v,1]) + K_j t, one approximate reciprocal, two pixels per instruction in packed fp32x2 arithmetic, floor on
//          the FMA pipe; accepted when the projection is a whole pixel inside the image and one per-pair line in (|d|, h_z)
//          bounds every forward-error margin below 1 px;
//   tier 2 (margin_pixel): the same evaluation with PER-PIXEL margins. mvcs_prepare derives a rigorous forward-error bound of
//          both evaluations against the real-valued result (gamma_n sum |a_i b_i| with n = 32 >= the 12-op chain, at u <= W,
//          v <= H): |dh_c| <= Bh_c d + Bt_c, which gives a margin on u_j, v_j, z_j; a pixel farther than the margin from
//          u_j in {0, W}, v_j in {0, H}, z_j = 0 and the z clamp provably takes the same mask decision on both paths;
//   tier 3 (exact_pixel, ~1 pixel in 10^4 at 504 x 504): the reference's operation order.
// Sampled values differ by a few ulp between tiers 1-2 and the reference (the tests allow 1e-5 on the per-pair MSE; the counts
// stay exact).
#include "common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>
#include <stdlib.h>

namespace vgpa {
namespace {

constexpr int MV_THREADS = 256;
constexpr int MV_PAIR_FLOATS = 64;  // exact path: invK(9) R(9) t(3) Kj(9) | fast path: M(9) Kt(3) Z(3) tz | margins: Au Cu Av Cv Az Cz Ahz Chz | interior line G1 G2
constexpr int MV_EXACT_FLOATS = 30;
constexpr int MV_FAST0 = 32;          // first fast-path constant
constexpr double MV_GAMMA = 32.0 * 5.9604644775390625e-08;   // 32 * 2^-24

__device__ bool inv3x3_d(const double* m, double* o) {
  const double c00 = m[4] * m[8] - m[5] * m[7];
  const double c01 = m[5] * m[6] - m[3] * m[8];
  const double c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  const double id = 1.0 / det;
  o[0] = c00 * id;
  o[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id;
  o[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id;
  o[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return det != 0.0;
}

// Gauss-Jordan with partial pivoting on [A | I], fp64.
__device__ void inv4x4_d(const double* a, double* inv) {
  double m[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) { m[r][c] = a[r * 4 + c]; m[r][4 + c] = (r == c) ? 1.0 : 0.0; }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    double best = fabs(m[col][col]);
    for (int r = col + 1; r < 4; ++r) { const double v = fabs(m[r][col]); if (v > best) { best = v; piv = r; } }
    if (piv != col)
      for (int c = 0; c < 8; ++c) { const double tmp = m[col][c]; m[col][c] = m[piv][c]; m[piv][c] = tmp; }
    const double d = 1.0 / m[col][col];
    for (int c = 0; c < 8; ++c) m[col][c] = m[col][c] * d;
    for (int r = 0; r < 4; ++r) {
      if (r == col) continue;
      const double f = m[r][col];
      for (int c = 0; c < 8; ++c) m[r][c] = m[r][c] - f * m[col][c];
    }
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) inv[r * 4 + c] = m[r][4 + c];
}

__global__ void mvcs_prepare_kernel(const float* __restrict__ Kmat, const float* __restrict__ Emat, int n_clips,
                                    int T, int k_dim, int e_rows, int H, int W, float* __restrict__ pairs) {
  const double Wd = static_cast<double>(W), Hd = static_cast<double>(H);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_pairs = n_clips * (T - 1);
  if (idx >= n_pairs) return;
  const int clip = idx / (T - 1), i = idx - clip * (T - 1), j = i + 1;
  const float* Ki = Kmat + (static_cast<long long>(clip) * T + i) * k_dim * k_dim;
  const float* Kj = Kmat + (static_cast<long long>(clip) * T + j) * k_dim * k_dim;
  const float* Ei = Emat + (static_cast<long long>(clip) * T + i) * e_rows * 4;
  const float* Ej = Emat + (static_cast<long long>(clip) * T + j) * e_rows * 4;
  double ki[9], kinv[9], ei[16], ej[16], einv[16];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) ki[r * 3 + c] = static_cast<double>(Ki[r * k_dim + c]);
  inv3x3_d(ki, kinv);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      ei[r * 4 + c] = (r < e_rows) ? static_cast<double>(Ei[r * 4 + c]) : (c == 3 ? 1.0 : 0.0);
      ej[r * 4 + c] = (r < e_rows) ? static_cast<double>(Ej[r * 4 + c]) : (c == 3 ? 1.0 : 0.0);
    }
  inv4x4_d(ei, einv);
  float* o = pairs + static_cast<long long>(idx) * MV_PAIR_FLOATS;
  for (int k = 0; k < 9; ++k) o[k] = static_cast<float>(kinv[k]);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 4; ++c) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s = s + ej[r * 4 + k] * einv[k * 4 + c];
      if (c < 3) o[9 + r * 3 + c] = static_cast<float>(s);
      else o[18 + r] = static_cast<float>(s);
    }
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o[21 + r * 3 + c] = Kj[r * k_dim + c];
  o[30] = 0.f; o[31] = 0.f;
  // ---- fast path: merged matrices from the SAME fp32-rounded factors the exact path multiplies, in fp64
  double kiv[9], R[9], t[3], kj[9];
  for (int k = 0; k < 9; ++k) { kiv[k] = static_cast<double>(o[k]); R[k] = static_cast<double>(o[9 + k]); kj[k] = static_cast<double>(o[21 + k]); }
  for (int k = 0; k < 3; ++k) t[k] = static_cast<double>(o[18 + k]);
  double RK[9], aRK[9], M[9], aM[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double s = 0.0, a = 0.0;
      for (int k = 0; k < 3; ++k) { s = s + R[r * 3 + k] * kiv[k * 3 + c]; a = a + fabs(R[r * 3 + k]) * fabs(kiv[k * 3 + c]); }
      RK[r * 3 + c] = s; aRK[r * 3 + c] = a;
    }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double s = 0.0, a = 0.0;
      for (int k = 0; k < 3; ++k) { s = s + kj[r * 3 + k] * RK[k * 3 + c]; a = a + fabs(kj[r * 3 + k]) * aRK[k * 3 + c]; }
      M[r * 3 + c] = s; aM[r * 3 + c] = a;
    }
  float* f = o + MV_FAST0;
  for (int k = 0; k < 9; ++k) f[k] = static_cast<float>(M[k]);
  double Bt[3], Bh[3];
  for (int r = 0; r < 3; ++r) {
    const double kt = kj[r * 3 + 0] * t[0] + kj[r * 3 + 1] * t[1] + kj[r * 3 + 2] * t[2];
    f[9 + r] = static_cast<float>(kt);
    Bt[r] = MV_GAMMA * (fabs(kj[r * 3 + 0] * t[0]) + fabs(kj[r * 3 + 1] * t[1]) + fabs(kj[r * 3 + 2] * t[2]));
    Bh[r] = MV_GAMMA * (aM[r * 3 + 0] * Wd + aM[r * 3 + 1] * Hd + aM[r * 3 + 2]);
  }
  for (int c = 0; c < 3; ++c) f[12 + c] = static_cast<float>(RK[6 + c]);      // z_j = d (RK[2,:] [u,v,1]) + t_z
  f[15] = static_cast<float>(t[2]);
  const double Bz = MV_GAMMA * (aRK[6] * Wd + aRK[7] * Hd + aRK[8]);
  // Error terms (rounded up). With dx = |hx_exact - hx_fast| <= Bh_x |d| + Bt_x and dz likewise, and the guard hz_fast > 4 dz:
  //   |u_exact - u_fast| <= (dx + |u_fast| dz) / (0.75 hz_fast) + |u_fast| 2^-21      (division / reciprocal rounding)
  // f[16..19] carry the 4/3; f[22..23] are the plain dz.
  const double up = 1.0 + 1e-6, k43 = 1.34;
  f[16] = static_cast<float>(Bh[0] * k43 * up);
  f[17] = static_cast<float>((Bt[0] + 1e-30) * k43 * up);
  f[18] = static_cast<float>(Bh[1] * k43 * up);
  f[19] = static_cast<float>((Bt[1] + 1e-30) * k43 * up);
  f[20] = static_cast<float>(Bz * up);
  f[21] = static_cast<float>((MV_GAMMA * fabs(t[2]) + 1e-30) * up);
  f[22] = static_cast<float>(Bh[2] * up);
  f[23] = static_cast<float>((Bt[2] + 1e-30) * up);
  // Interior test of mvcs_pairs_kernel: h_z > G1 |d| + G2 must imply (i) margin_pixel's guard h_z > 4 e_z + 2e-8 and (ii) its
  // u / v margins below 1 px for |u_j| <= W, |v_j| <= H:  (W 1.34 e_z + f16 |d| + f17) rz + W 2^-21 < 0.99 with
  // rz <= (1 + 2^-22) / h_z. Built from the rounded f[] the margin path itself uses; kappa = 0.95 leaves room for W 2^-21
  // (W <= 16384), the reciprocal's error and the rounding of the line's own fp32 evaluation.
  {
    const double f16 = f[16], f17 = f[17], f18 = f[18], f19 = f[19], f22 = f[22], f23 = f[23], kappa = 0.95;
    const double s1 = fmax(Wd * 1.34 * f22 + f16, Hd * 1.34 * f22 + f18) / kappa;
    const double s2 = fmax(Wd * 1.34 * f23 + f17, Hd * 1.34 * f23 + f19) / kappa;
    double G1 = fmax(s1, 4.0 * f22) * up, G2 = fmax(s2, 4.0 * f23 + 2e-8) * up;
    if (W > 16384 || H > 16384 || !(G1 == G1) || !(G2 == G2)) { G1 = 0.0; G2 = INFINITY; }   // no pixel passes: margin path only
    f[24] = static_cast<float>(G1 * up);
    f[25] = static_cast<float>(G2 * up);
  }
}

__device__ __forceinline__ float fetch_zero_pad(const float* __restrict__ img, int x, int y, int W, int H) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(img + static_cast<long long>(y) * W + x) : 0.0f;
}

// One pixel in the reference's fp32 operation order (mvcs.py:64-99; no FMA contraction, IEEE division). Returns the mask
// decision and, when inside, the squared error. Out of line: about one pixel in 10^4 comes here.
__device__ __noinline__ bool exact_pixel(const float* __restrict__ sp, const float* __restrict__ dj, int px, int py, float d,
                                         int W, int H, float* e2) {
  const float Wm1 = static_cast<float>(W - 1), Hm1 = static_cast<float>(H - 1);
  const float Wf = static_cast<float>(W), Hf = static_cast<float>(H);
  const float u = static_cast<float>(px), v = static_cast<float>(py);
  // p_i = (K_i^-1 @ [u, v, 1]) * d          (mvcs.py:64-66)
  const float xi = ((sp[0] * u + sp[1] * v) + sp[2]) * d;
  const float yi = ((sp[3] * u + sp[4] * v) + sp[5]) * d;
  const float zi = ((sp[6] * u + sp[7] * v) + sp[8]) * d;
  // p_j = R @ p_i + t                       (mvcs.py:70-72)
  const float xj = ((sp[9] * xi + sp[10] * yi) + sp[11] * zi) + sp[18];
  const float yj = ((sp[12] * xi + sp[13] * yi) + sp[14] * zi) + sp[19];
  const float zj = ((sp[15] * xi + sp[16] * yi) + sp[17] * zi) + sp[20];
  // homogeneous projection with K_j, clamp z  (mvcs.py:75-81)
  const float hx = (sp[21] * xj + sp[22] * yj) + sp[23] * zj;
  const float hy = (sp[24] * xj + sp[25] * yj) + sp[26] * zj;
  const float hz = (sp[27] * xj + sp[28] * yj) + sp[29] * zj;
  const float zc = fmaxf(hz, 1e-8f);
  const float uj = hx / zc, vj = hy / zc;
  // normalise to [-1, 1] and back (grid_sample, align_corners=True)   (mvcs.py:85-95)
  const float gu = (2.0f * uj) / Wm1 - 1.0f;
  const float gv = (2.0f * vj) / Hm1 - 1.0f;
  const float ix = ((gu + 1.0f) / 2.0f) * Wm1;
  const float iy = ((gv + 1.0f) / 2.0f) * Hm1;
  const bool in_mask = (uj >= 0.0f) && (uj < Wf) && (vj >= 0.0f) && (vj < Hf) && (zj > 0.0f);  // mvcs.py:99
  if (!in_mask) return false;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float tx = ix - fx, ty = iy - fy;
  const float w_nw = (1.0f - tx) * (1.0f - ty), w_ne = tx * (1.0f - ty);
  const float w_sw = (1.0f - tx) * ty, w_se = tx * ty;
  const float s = ((fetch_zero_pad(dj, x0, y0, W, H) * w_nw + fetch_zero_pad(dj, x0 + 1, y0, W, H) * w_ne) +
                   fetch_zero_pad(dj, x0, y0 + 1, W, H) * w_sw) + fetch_zero_pad(dj, x0 + 1, y0 + 1, W, H) * w_se;
  const float e = s - zj;
  *e2 = e * e;
  return true;
}

__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2: two pixels per FMA-pipe instruction; explicit PTX, so the file's
//      --fmad=false does not touch it)
__device__ __forceinline__ uint64_t p2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t p2(float c) { return p2(c, c); }
__device__ __forceinline__ void u2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add_rm2(uint64_t a, uint64_t b) {   // round toward -inf
  uint64_t d;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// TIER 2 (and 3): one pixel with per-pixel forward-error margins; a pixel that is not provably decided by the merged-matrix
// evaluation is handed to `exact_pixel`. Out of line: only pixels that fail the interior test of the kernel below come here
// (the 1-pixel border ring, projections outside the image, non-positive depth: a few percent of a typical clip).
__device__ __noinline__ bool margin_pixel(const float* __restrict__ prec, const float* __restrict__ dj, int px, int py, float d,
                                          int W, int H, float* e2) {
  float fp[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) fp[k] = __ldg(prec + MV_FAST0 + k);
  const float Wf = static_cast<float>(W), Hf = static_cast<float>(H);
  const float half_w = 0.5f * Wf, half_h = 0.5f * Hf;
  const float u = static_cast<float>(px), v = static_cast<float>(py);
  const float mx = __fmaf_rn(fp[1], v, fp[2]), my = __fmaf_rn(fp[4], v, fp[5]), mzr = __fmaf_rn(fp[7], v, fp[8]);
  const float mzz = __fmaf_rn(fp[13], v, fp[14]);
  // h = d (M [u,v,1]) + K_j t ;  z_j = d (Z [u,v,1]) + t_z
  const float hx = __fmaf_rn(d, __fmaf_rn(fp[0], u, mx), fp[9]);
  const float hy = __fmaf_rn(d, __fmaf_rn(fp[3], u, my), fp[10]);
  const float hz = __fmaf_rn(d, __fmaf_rn(fp[6], u, mzr), fp[11]);
  const float zj = __fmaf_rn(d, __fmaf_rn(fp[12], u, mzz), fp[15]);
  const float rz = rcp_fast(fmaxf(hz, 1e-8f));
  const float uj = hx * rz, vj = hy * rz;
  // signed distance to the nearest mask boundary (positive inside) against the forward-error margins
  const float su = half_w - fabsf(uj - half_w), sv = half_h - fabsf(vj - half_h);
  const float ad = fabsf(d), auj = fabsf(uj), avj = fabsf(vj);
  const float ez = __fmaf_rn(fp[22], ad, fp[23]), ezs = 1.34f * ez;
  const float mu = __fmaf_rn(__fmaf_rn(auj, ezs, __fmaf_rn(fp[16], ad, fp[17])), rz, auj * 4.76837158203125e-07f);
  const float mv = __fmaf_rn(__fmaf_rn(avj, ezs, __fmaf_rn(fp[18], ad, fp[19])), rz, avj * 4.76837158203125e-07f);
  const float mz = __fmaf_rn(fp[20], ad, fp[21]);
  // written so that a NaN anywhere selects the exact path; hz > 4 ez + 2e-8 also keeps both paths off the 1e-8 clamp;
  // the 1e-3 absolute slack covers the rounding of su / sv themselves (|u_j - W/2| is rounded once: <= W 2^-24)
  const bool clear = (fabsf(su) > mu + 1e-3f) && (fabsf(sv) > mv + 1e-3f) && (fabsf(zj) > mz) && (hz > __fmaf_rn(4.0f, ez, 2e-8f));
  if (!clear) {                                            // TIER 3: the reference's operation order
    float spx[MV_EXACT_FLOATS];
#pragma unroll
    for (int q = 0; q < MV_EXACT_FLOATS; ++q) spx[q] = __ldg(prec + q);
    return exact_pixel(spx, dj, px, py, d, W, H, e2);
  }
  if (!((su > 0.0f) && (sv > 0.0f) && (zj > 0.0f))) return false;
  // grid_sample(align_corners=True) maps the normalised coordinate back to u_j, v_j (mvcs.py:85-95; the reference's round
  // trip adds ~1e-5 px of rounding noise, which this path does not reproduce)
  const float fx = floorf(uj), fy = floorf(vj);
  // u_j in [0, W), v_j in [0, H): the base texel is inside; only the +1 neighbours can leave the image (zero padding)
  const int x0 = min(static_cast<int>(fx), W - 1), y0 = min(static_cast<int>(fy), H - 1);
  const float tx = uj - fx, ty = vj - fy;
  const bool okx = x0 + 1 < W, oky = y0 + 1 < H;
  const float* r0 = dj + (static_cast<long long>(y0) * W + x0);
  const float v00 = __ldg(r0);
  const float a01 = okx ? __ldg(r0 + 1) : 0.f, a10 = oky ? __ldg(r0 + W) : 0.f, a11 = (okx && oky) ? __ldg(r0 + W + 1) : 0.f;
  const float top = __fmaf_rn(tx, a01 - v00, v00), bot = __fmaf_rn(tx, a11 - a10, a10);
  const float e = __fmaf_rn(ty, bot - top, top) - zj;
  *e2 = e * e;
  return true;
}

// TIER 1: one warp walks one image row; a thread takes the pixel pair (x, x + 32) and evaluates both in packed fp32x2
// arithmetic. A pixel is finished here when it is provably in the mask by a whole pixel:
//   * floor(u_j) in [1, W - 2] and floor(v_j) in [1, H - 2] (so u_j, v_j are >= 1 px from every mask boundary, and the four
//     bilinear taps are inside the image: no zero padding, no clamping), with floor() taken on the FMA pipe by a
//     round-toward-minus-infinity add of 1.5 * 2^23 (exact for 0 <= x < 2^22; any other input, NaN included, yields an
//     integer outside the accepted range);
//   * h_z > G1 |d| + G2, the per-pair line (mvcs_prepare) that bounds BOTH margin_pixel's h_z guard (4 e_z + 2e-8) and
//     (W 1.34 e_z + Bu |d| + Cu) / 0.95, the numerator of its u / v margin at |u_j| <= W: margin < 0.95 + W 2^-21 < 1 px;
//   * z_j > Az |d| + Cz, margin_pixel's own z margin (>= 0, so z_j > 0).
// These imply margin_pixel's `clear && in_mask`, so the pixel takes the same mask decision as the reference, and the sampled
// value is computed by the same formula as in margin_pixel. Everything else goes to margin_pixel.
// 4 CTAs (32 warps) per SM at 64 registers: measured 0.956 ms per 128 clips of 10 x 504^2 against 1.119 ms at 3 CTAs / 80 registers.
__global__ void __launch_bounds__(MV_THREADS, 4)
mvcs_pairs_kernel(const float* __restrict__ depths, const float* __restrict__ pairs, int T, int H, int W,
                  int blocks_per_pair, double* __restrict__ part_sum, unsigned int* __restrict__ part_cnt) {
  const int pair_i = blockIdx.y;        // 0..T-2
  const int clip = blockIdx.z;
  const long long pair_idx = static_cast<long long>(clip) * (T - 1) + pair_i;
  const float* prec = pairs + pair_idx * MV_PAIR_FLOATS;
  __shared__ float s_fp[32];            // the pair's fast-path constants; the row-start ones are re-read from here per row
  if (threadIdx.x < 32) s_fp[threadIdx.x] = __ldg(prec + MV_FAST0 + threadIdx.x);
  __syncthreads();
  const float* di = depths + (static_cast<long long>(clip) * T + pair_i) * H * W;
  const float* dj = di + static_cast<long long>(H) * W;
  // keep the gather base in a register pair: under the 64-register cap the compiler otherwise re-derives it from the kernel
  // parameters inside the loop (a dozen 64-bit integer instructions per pixel pair)
  asm volatile("" : "+l"(dj));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // constants of the inner loop (scalar registers: the packed instructions broadcast a 32-bit operand to both halves)
  const float m0 = s_fp[0], m3 = s_fp[3], m6 = s_fp[6], z0 = s_fp[12];
  const float kt0 = s_fp[9], kt1 = s_fp[10], kt2 = s_fp[11], tz = s_fp[15];
  const float az = s_fp[20], cz = s_fp[21], g1 = s_fp[24], g2 = s_fp[25];
  const unsigned int xr = static_cast<unsigned int>(W - 2), yr = static_cast<unsigned int>(H - 2), Wu = static_cast<unsigned int>(W);
  double acc = 0.0;
  unsigned int cnt = 0;
  for (int row = blockIdx.x * (MV_THREADS / 32) + warp; row < H; row += blocks_per_pair * (MV_THREADS / 32)) {
    const float v = static_cast<float>(row);
    const float rx = __fmaf_rn(s_fp[1], v, s_fp[2]), ry = __fmaf_rn(s_fp[4], v, s_fp[5]), rzr = __fmaf_rn(s_fp[7], v, s_fp[8]);
    const float rzz = __fmaf_rn(s_fp[13], v, s_fp[14]);
    const float* drow = di + static_cast<long long>(row) * W;
    float part_a = 0.f, part_b = 0.f;                           // fp32 sums of this thread's squared errors in this row
    // depth_i of the NEXT pixel pair is requested before the current pair's arithmetic: the HBM latency of the streamed
    // operand overlaps the L2 latency of the gathers instead of adding to it
    float da = (lane < W) ? __ldg(drow + lane) : 0.f;
    float db = (lane + 32 < W) ? __ldg(drow + lane + 32) : 0.f;
#pragma unroll 1
    for (int x = lane; x < W; x += 64) {
      const bool live_b = x + 32 < W;
      const float da_n = (x + 64 < W) ? __ldg(drow + x + 64) : 0.f;
      const float db_n = (x + 96 < W) ? __ldg(drow + x + 96) : 0.f;
      const float ua = static_cast<float>(x);
      const uint64_t uu = p2(ua, ua + 32.0f), dd = p2(da, db);
      // h = d (M [u,v,1]) + K_j t ;  z_j = d (Z [u,v,1]) + t_z
      const uint64_t hx = fma2(dd, fma2(p2(m0), uu, p2(rx)), p2(kt0));
      const uint64_t hy = fma2(dd, fma2(p2(m3), uu, p2(ry)), p2(kt1));
      const uint64_t hz = fma2(dd, fma2(p2(m6), uu, p2(rzr)), p2(kt2));
      const uint64_t zj = fma2(dd, fma2(p2(z0), uu, p2(rzz)), p2(tz));
      float hza, hzb;
      u2(hz, hza, hzb);
      const uint64_t rz = p2(rcp_fast(fmaxf(hza, 1e-8f)), rcp_fast(fmaxf(hzb, 1e-8f)));
      const uint64_t uj = mul2(hx, rz), vj = mul2(hy, rz);
      const uint64_t magic = p2(12582912.0f);                   // 1.5 * 2^23: ulp 1, so a round-down add leaves floor(x)
      const uint64_t tu = add_rm2(uj, magic), tv = add_rm2(vj, magic);
      const uint64_t tx = sub2(uj, sub2(tu, magic)), ty = sub2(vj, sub2(tv, magic));
      const uint64_t ad = dd & 0x7fffffff7fffffffull;
      const uint64_t gz = fma2(p2(g1), ad, p2(g2)), mz = fma2(p2(az), ad, p2(cz));
      float tua, tub, tva, tvb, gza, gzb, mza, mzb, zja, zjb;
      u2(tu, tua, tub); u2(tv, tva, tvb); u2(gz, gza, gzb); u2(mz, mza, mzb); u2(zj, zja, zjb);
      const int x0a = __float_as_int(tua) - 0x4B400000, x0b = __float_as_int(tub) - 0x4B400000;
      const int y0a = __float_as_int(tva) - 0x4B400000, y0b = __float_as_int(tvb) - 0x4B400000;
      const bool in_a = (static_cast<unsigned int>(x0a - 1) < xr) && (static_cast<unsigned int>(y0a - 1) < yr) && (hza > gza) && (zja > mza);
      const bool in_b = live_b && (static_cast<unsigned int>(x0b - 1) < xr) && (static_cast<unsigned int>(y0b - 1) < yr) && (hzb > gzb) &&
                        (zjb > mzb);
      // the 8 gathers of depth_j: addresses forced valid (offset 0) for pixels that leave this path
      // (unsigned 32-bit element offsets: one IMAD.WIDE.U32 per row pointer)
      const unsigned int oa = in_a ? static_cast<unsigned int>(y0a * W + x0a) : 0u;
      const unsigned int ob = in_b ? static_cast<unsigned int>(y0b * W + x0b) : 0u;
      const float* ra0 = dj + oa;
      const float* ra1 = dj + (oa + Wu);
      const float* rb0 = dj + ob;
      const float* rb1 = dj + (ob + Wu);
      const float a00 = __ldg(ra0), a01 = __ldg(ra0 + 1), a10 = __ldg(ra1), a11 = __ldg(ra1 + 1);
      const float b00 = __ldg(rb0), b01 = __ldg(rb0 + 1), b10 = __ldg(rb1), b11 = __ldg(rb1 + 1);
      const uint64_t v00 = p2(a00, b00), v10 = p2(a10, b10);
      const uint64_t top = fma2(tx, sub2(p2(a01, b01), v00), v00), bot = fma2(tx, sub2(p2(a11, b11), v10), v10);
      float ea, eb;
      u2(sub2(fma2(ty, sub2(bot, top), top), zj), ea, eb);
      part_a = in_a ? __fmaf_rn(ea, ea, part_a) : part_a;
      part_b = in_b ? __fmaf_rn(eb, eb, part_b) : part_b;
      cnt += (in_a ? 1u : 0u) + (in_b ? 1u : 0u);
      if (!in_a || (live_b && !in_b)) {                         // border ring, outside the image, z <= 0, NaN: margin / exact path
        float e2;
        if (!in_a && margin_pixel(prec, dj, x, row, da, W, H, &e2)) { acc += static_cast<double>(e2); ++cnt; }
        if (live_b && !in_b && margin_pixel(prec, dj, x + 32, row, db, W, H, &e2)) { acc += static_cast<double>(e2); ++cnt; }
      }
      da = da_n; db = db_n;
    }
    acc += static_cast<double>(part_a) + static_cast<double>(part_b);
  }
  // block reduction in a fixed order
  __shared__ double s_sum[MV_THREADS / 32];
  __shared__ unsigned int s_cnt[MV_THREADS / 32];
  acc = warp_sum_d(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) { s_sum[warp] = acc; s_cnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0; unsigned int c = 0;
    for (int w = 0; w < MV_THREADS / 32; ++w) { t += s_sum[w]; c += s_cnt[w]; }
    part_sum[pair_idx * blocks_per_pair + blockIdx.x] = t;
    part_cnt[pair_idx * blocks_per_pair + blockIdx.x] = c;
  }
}

__global__ void mvcs_finalize_kernel(const double* __restrict__ part_sum, const unsigned int* __restrict__ part_cnt,
                                     int n_clips, int T, int blocks_per_pair, double* __restrict__ pair_mse,
                                     long long* __restrict__ pair_cnt, double* __restrict__ scores) {
  const int clip = blockIdx.x * blockDim.x + threadIdx.x;
  if (clip >= n_clips) return;
  double total = 0.0;
  int used = 0;
  for (int i = 0; i < T - 1; ++i) {
    const long long p = static_cast<long long>(clip) * (T - 1) + i;
    double s = 0.0; long long c = 0;
    for (int b = 0; b < blocks_per_pair; ++b) { s += part_sum[p * blocks_per_pair + b]; c += part_cnt[p * blocks_per_pair + b]; }
    const double mse = c > 0 ? s / static_cast<double>(c) : 0.0;
    if (pair_mse) pair_mse[p] = mse;
    if (pair_cnt) pair_cnt[p] = c;
    if (c > 0) { total += mse; ++used; }      // empty-mask pairs are skipped, not counted (mvcs.py:101-104)
  }
  scores[clip] = used > 0 ? exp(-(total / static_cast<double>(used))) : 0.0;   // mvcs.py:108-113
}

}  // namespace
}  // namespace vgpa

extern "C" size_t vgpa_mvcs_workspace_bytes(int n_clips, int T, int H, int W) {
  if (n_clips <= 0 || T <= 1 || H <= 0 || W <= 0) return 256;
  const long long n_pairs = static_cast<long long>(n_clips) * (T - 1);
  const int bpp = vgpa_mvcs_blocks_per_pair(n_clips, T, H, W);
  return static_cast<size_t>(n_pairs) * (vgpa::MV_PAIR_FLOATS * 4 + static_cast<size_t>(bpp) * 16) + 256;
}

extern "C" int vgpa_mvcs_blocks_per_pair(int n_clips, int T, int H, int W) {
  (void)W;
  const long long rows_per_pass = vgpa::MV_THREADS / 32;           // one warp walks one image row at a time
  long long need = (H + rows_per_pass - 1) / rows_per_pass;
  // enough blocks to fill the machine (148 SMs x 8 resident CTAs) without shrinking below one pass per block
  const long long n_pairs = static_cast<long long>(n_clips > 0 ? n_clips : 1) * (T > 1 ? T - 1 : 1);
  long long want = (148LL * 8 + n_pairs - 1) / n_pairs;
  if (want < 1) want = 1;
  if (need > want) need = want;
  if (need < 1) need = 1;
  return static_cast<int>(need);
}

extern "C" int vgpa_mvcs_batch(const float* d_depths, const float* d_intrinsics, const float* d_extrinsics, int n_clips,
                               int T, int H, int W, int k_dim, int e_rows, void* d_workspace, size_t workspace_bytes,
                               double* d_pair_mse, int64_t* d_pair_cnt, double* d_scores, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(d_depths && d_intrinsics && d_extrinsics && d_scores, "vgpa_mvcs_batch: null pointer");
  VGPA_CHECK(n_clips > 0 && T >= 1 && H > 0 && W > 0, "vgpa_mvcs_batch: bad shape clips=%d T=%d H=%d W=%d", n_clips, T, H, W);
  VGPA_CHECK(k_dim == 3 || k_dim == 4, "vgpa_mvcs_batch: intrinsics must be 3x3 or 4x4 (k_dim=%d)", k_dim);
  VGPA_CHECK(e_rows == 3 || e_rows == 4, "vgpa_mvcs_batch: extrinsics must be 3x4 or 4x4 (e_rows=%d)", e_rows);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (T == 1) {  // no pair: the reference returns 0.0 (mvcs.py:108-109)
    VGPA_CUDA(cudaMemsetAsync(d_scores, 0, sizeof(double) * n_clips, s));
    return 0;
  }
  VGPA_CHECK(d_workspace != nullptr && workspace_bytes >= vgpa_mvcs_workspace_bytes(n_clips, T, H, W),
             "vgpa_mvcs_batch: workspace too small (%zu < %zu)", workspace_bytes, vgpa_mvcs_workspace_bytes(n_clips, T, H, W));
  VGPA_CHECK((reinterpret_cast<uintptr_t>(d_depths) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0,
             "vgpa_mvcs_batch: depths must be 16-byte and workspace 256-byte aligned");
  const int n_pairs = n_clips * (T - 1);
  const int bpp = vgpa_mvcs_blocks_per_pair(n_clips, T, H, W);
  uint8_t* ws = static_cast<uint8_t*>(d_workspace);
  double* part_sum = reinterpret_cast<double*>(ws);
  unsigned int* part_cnt = reinterpret_cast<unsigned int*>(ws + static_cast<size_t>(n_pairs) * bpp * 8);
  float* pairs = reinterpret_cast<float*>(ws + static_cast<size_t>(n_pairs) * bpp * 16);
  mvcs_prepare_kernel<<<(n_pairs + 127) / 128, 128, 0, s>>>(d_intrinsics, d_extrinsics, n_clips, T, k_dim, e_rows, H, W, pairs);
  VGPA_LAUNCH_CHECK("mvcs_prepare_kernel");
  VGPA_CHECK(n_clips <= 65535 && T - 1 <= 65535, "vgpa_mvcs_batch: too many clips per launch (%d)", n_clips);
  VGPA_CHECK(static_cast<long long>(H) * W < (1LL << 24) && H >= 2 && W >= 2, "vgpa_mvcs_batch: frames must be between 2x2 and 2^24 pixels (H=%d W=%d)", H, W);
  dim3 grid(bpp, T - 1, n_clips);
  mvcs_pairs_kernel<<<grid, MV_THREADS, 0, s>>>(d_depths, pairs, T, H, W, bpp, part_sum, part_cnt);
  VGPA_LAUNCH_CHECK("mvcs_pairs_kernel");
  mvcs_finalize_kernel<<<(n_clips + 127) / 128, 128, 0, s>>>(part_sum, part_cnt, n_clips, T, bpp, d_pair_mse,
                                                            reinterpret_cast<long long*>(d_pair_cnt), d_scores);
  VGPA_LAUNCH_CHECK("mvcs_finalize_kernel");
  return 0;
}
