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
// --fmad=false, IEEE division). Evaluated for every pixel that chain costs ~280 instructions and made the kernel
// issue-bound at 0.15 of the HBM roofline. Only pixels whose projection lands within rounding distance of a mask boundary
// can decide differently under another evaluation order, so every pixel first takes a FAST path: the three matrices are
// merged per pair (M = K_j R K_i^-1 in fp64, rounded once), h = d (M [u,v,1]) + K_j t with explicit FMAs, one approximate
// reciprocal. mvcs_prepare also derives a rigorous forward-error bound of BOTH evaluations against the real-valued result
// (gamma_n sum |a_i b_i| with n = 32 >= the 12-op chain, at u <= W, v <= H): |dh_c| <= Bh_c d + Bt_c, which gives a per-pixel
// margin on u_j, v_j, z_j. A pixel closer than the margin to u_j in {0, W}, v_j in {0, H}, z_j = 0 or the z clamp is
// re-evaluated by `exact_pixel` (the reference's operation order); all other pixels provably take the same mask decision on
// both paths. Sampled values differ by a few ulp between the paths (the tests allow 1e-5 on the per-pair MSE; the counts
// stay exact). At 504 x 504 about 1 pixel in 10^4 takes the exact path.
#include "common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {

constexpr int MV_THREADS = 256;
constexpr int MV_PIX_PER_THREAD = 4;
constexpr int MV_PAIR_FLOATS = 64;  // exact path: invK(9) R(9) t(3) Kj(9) | fast path: M(9) Kt(3) Z(3) tz | margins: Au Cu Av Cv Az Cz Ahz Chz
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

// ONE_ROW: W % 4 == 0, so the 4 consecutive pixels of a thread share their row (the row terms are hoisted) and the depth_i
// vector load is always in range.
template <bool ONE_ROW>
__global__ void __launch_bounds__(MV_THREADS, 3)
mvcs_pairs_kernel(const float* __restrict__ depths, const float* __restrict__ pairs, int T, int H, int W,
                  int blocks_per_pair, double* __restrict__ part_sum, unsigned int* __restrict__ part_cnt) {
  const int pair_i = blockIdx.y;        // 0..T-2
  const int clip = blockIdx.z;
  const long long pair_idx = static_cast<long long>(clip) * (T - 1) + pair_i;
  const float* prec = pairs + pair_idx * MV_PAIR_FLOATS;
  float fp[24];                         // fast-path constants in registers
#pragma unroll
  for (int k = 0; k < 24; ++k) fp[k] = __ldg(prec + MV_FAST0 + k);
  const float ez_scale = 1.34f;
  const float* di = depths + (static_cast<long long>(clip) * T + pair_i) * H * W;
  const float* dj = di + static_cast<long long>(H) * W;
  const int HW = H * W;
  const float Wf = static_cast<float>(W), Hf = static_cast<float>(H);
  const float half_w = 0.5f * Wf, half_h = 0.5f * Hf;
  const float inv_w = 1.0f / Wf;
  double acc = 0.0;
  unsigned int cnt = 0;
  const int stride = blocks_per_pair * MV_THREADS * MV_PIX_PER_THREAD;
  for (int base = (blockIdx.x * MV_THREADS + threadIdx.x) * MV_PIX_PER_THREAD; base < HW; base += stride) {
    float d0, d1, d2, d3;
    if (ONE_ROW) {
      const float4 q = *reinterpret_cast<const float4*>(di + base);
      d0 = q.x; d1 = q.y; d2 = q.z; d3 = q.w;
    } else {
      d0 = di[base];
      d1 = (base + 1 < HW) ? di[base + 1] : 0.f;
      d2 = (base + 2 < HW) ? di[base + 2] : 0.f;
      d3 = (base + 3 < HW) ? di[base + 3] : 0.f;
    }
    const float dd[MV_PIX_PER_THREAD] = {d0, d1, d2, d3};
    // row / column of the first pixel without an integer division (exact below 2^24 pixels; checked by the host wrapper)
    int py0 = __float2int_rz(__int2float_rz(base) * inv_w);
    int px0 = base - py0 * W;
    if (px0 >= W) { px0 -= W; ++py0; }
    if (px0 < 0) { px0 += W; --py0; }
    const float u0 = static_cast<float>(px0), v0 = static_cast<float>(py0);
    // row terms of M [u,v,1] and Z [u,v,1]
    const float rx = __fmaf_rn(fp[1], v0, fp[2]), ry = __fmaf_rn(fp[4], v0, fp[5]), rzr = __fmaf_rn(fp[7], v0, fp[8]);
    const float rzz = __fmaf_rn(fp[13], v0, fp[14]);
    // ---- phase 1: projection, mask and sample coordinates of the 4 pixels (no loads, no branches)
    float zj[MV_PIX_PER_THREAD], tx[MV_PIX_PER_THREAD], ty[MV_PIX_PER_THREAD];
    int off[MV_PIX_PER_THREAD];
    bool take[MV_PIX_PER_THREAD], redo[MV_PIX_PER_THREAD], okx[MV_PIX_PER_THREAD], oky[MV_PIX_PER_THREAD];
#pragma unroll
    for (int k = 0; k < MV_PIX_PER_THREAD; ++k) {
      float u = u0 + static_cast<float>(k), mx = rx, my = ry, mzr = rzr, mzz = rzz;
      bool live = true;
      if (!ONE_ROW) {
        int px = px0 + k, py = py0;
        if (px >= W) { px -= W; ++py; }
        live = base + k < HW;
        u = static_cast<float>(px);
        const float v = static_cast<float>(py);
        mx = __fmaf_rn(fp[1], v, fp[2]); my = __fmaf_rn(fp[4], v, fp[5]); mzr = __fmaf_rn(fp[7], v, fp[8]); mzz = __fmaf_rn(fp[13], v, fp[14]);
      }
      const float d = dd[k];
      // h = d (M [u,v,1]) + K_j t ;  z_j = d (Z [u,v,1]) + t_z
      const float hx = __fmaf_rn(d, __fmaf_rn(fp[0], u, mx), fp[9]);
      const float hy = __fmaf_rn(d, __fmaf_rn(fp[3], u, my), fp[10]);
      const float hz = __fmaf_rn(d, __fmaf_rn(fp[6], u, mzr), fp[11]);
      zj[k] = __fmaf_rn(d, __fmaf_rn(fp[12], u, mzz), fp[15]);
      const float rz = rcp_fast(fmaxf(hz, 1e-8f));
      const float uj = hx * rz, vj = hy * rz;
      // signed distance to the nearest mask boundary (positive inside) against the forward-error margins
      const float su = half_w - fabsf(uj - half_w), sv = half_h - fabsf(vj - half_h);
      const float ad = fabsf(d), auj = fabsf(uj), avj = fabsf(vj);
      const float ez = __fmaf_rn(fp[22], ad, fp[23]), ezs = ez_scale * ez;
      const float mu = __fmaf_rn(__fmaf_rn(auj, ezs, __fmaf_rn(fp[16], ad, fp[17])), rz, auj * 4.76837158203125e-07f);
      const float mv = __fmaf_rn(__fmaf_rn(avj, ezs, __fmaf_rn(fp[18], ad, fp[19])), rz, avj * 4.76837158203125e-07f);
      const float mz = __fmaf_rn(fp[20], ad, fp[21]);
      // written so that a NaN anywhere selects the exact path; hz > 4 ez + 2e-8 also keeps both paths off the 1e-8 clamp;
      // the 1e-3 absolute slack covers the rounding of su / sv themselves (|u_j - W/2| is rounded once: <= W 2^-24)
      const bool clear = (fabsf(su) > mu + 1e-3f) && (fabsf(sv) > mv + 1e-3f) && (fabsf(zj[k]) > mz) && (hz > __fmaf_rn(4.0f, ez, 2e-8f));
      const bool in_mask = (su > 0.0f) && (sv > 0.0f) && (zj[k] > 0.0f);
      take[k] = live && clear && in_mask;
      redo[k] = live && !clear;
      // grid_sample(align_corners=True) maps the normalised coordinate back to u_j, v_j (mvcs.py:85-95; the reference's round
      // trip adds ~1e-5 px of rounding noise, which this path does not reproduce)
      const float ix = take[k] ? uj : 0.0f, iy = take[k] ? vj : 0.0f;
      const float fx = floorf(ix), fy = floorf(iy);
      // u_j in [0, W), v_j in [0, H): the base texel is inside; only the +1 neighbours can leave the image (zero padding)
      const int x0 = min(static_cast<int>(fx), W - 1), y0 = min(static_cast<int>(fy), H - 1);
      tx[k] = ix - fx; ty[k] = iy - fy;
      okx[k] = x0 + 1 < W; oky[k] = y0 + 1 < H;
      off[k] = y0 * W + x0;
    }
    // ---- phase 2: the 16 gathers of depth_j, all addresses valid (clamped), issued together
    float v00[MV_PIX_PER_THREAD], v01[MV_PIX_PER_THREAD], v10[MV_PIX_PER_THREAD], v11[MV_PIX_PER_THREAD];
#pragma unroll
    for (int k = 0; k < MV_PIX_PER_THREAD; ++k) {
      const int ox = okx[k] ? 1 : 0, oy = oky[k] ? W : 0;
      const float* r0 = dj + off[k];
      v00[k] = __ldg(r0); v01[k] = __ldg(r0 + ox); v10[k] = __ldg(r0 + oy); v11[k] = __ldg(r0 + oy + ox);
    }
    // ---- phase 3: bilinear sample, squared error
    float part = 0.f;                                      // fp32 sum of this thread's 4 squared errors, added in fp64 below
#pragma unroll
    for (int k = 0; k < MV_PIX_PER_THREAD; ++k) {
      const float a01 = okx[k] ? v01[k] : 0.f, a10 = oky[k] ? v10[k] : 0.f, a11 = (okx[k] && oky[k]) ? v11[k] : 0.f;
      const float top = __fmaf_rn(tx[k], a01 - v00[k], v00[k]), bot = __fmaf_rn(tx[k], a11 - a10, a10);
      const float e = __fmaf_rn(ty[k], bot - top, top) - zj[k];
      part = take[k] ? __fmaf_rn(e, e, part) : part;
      cnt += take[k] ? 1u : 0u;
    }
    acc += static_cast<double>(part);
    // ---- phase 4 (rare): pixels too close to a mask boundary are re-evaluated in the reference's operation order
    if (redo[0] || redo[1] || redo[2] || redo[3]) {
      float spx[MV_EXACT_FLOATS];
#pragma unroll
      for (int q = 0; q < MV_EXACT_FLOATS; ++q) spx[q] = __ldg(prec + q);
#pragma unroll
      for (int k = 0; k < MV_PIX_PER_THREAD; ++k) {
        if (!redo[k]) continue;
        int py = py0, px = px0 + k;
        if (px >= W) { px -= W; ++py; }
        float e2;
        if (exact_pixel(spx, dj, px, py, dd[k], W, H, &e2)) { acc += static_cast<double>(e2); ++cnt; }
      }
    }
  }
  // block reduction in a fixed order
  __shared__ double s_sum[MV_THREADS / 32];
  __shared__ unsigned int s_cnt[MV_THREADS / 32];
  acc = warp_sum_d(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  const long long hw = static_cast<long long>(H) * W;
  const long long per_block = vgpa::MV_THREADS * vgpa::MV_PIX_PER_THREAD;
  long long need = (hw + per_block - 1) / per_block;
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
  if ((W & 3) == 0) mvcs_pairs_kernel<true><<<grid, MV_THREADS, 0, s>>>(d_depths, pairs, T, H, W, bpp, part_sum, part_cnt);
  else mvcs_pairs_kernel<false><<<grid, MV_THREADS, 0, s>>>(d_depths, pairs, T, H, W, bpp, part_sum, part_cnt);
  VGPA_LAUNCH_CHECK("mvcs_pairs_kernel");
  mvcs_finalize_kernel<<<(n_clips + 127) / 128, 128, 0, s>>>(part_sum, part_cnt, n_clips, T, bpp, d_pair_mse,
                                                            reinterpret_cast<long long*>(d_pair_cnt), d_scores);
  VGPA_LAUNCH_CHECK("mvcs_finalize_kernel");
  return 0;
}
