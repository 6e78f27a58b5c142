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
// The per-pixel arithmetic keeps the reference's operation order in fp32 and this file is compiled
// with --fmad=false, so mask decisions match the numpy oracle bit for bit.
#include "common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {

constexpr int MV_THREADS = 256;
constexpr int MV_PIX_PER_THREAD = 4;
constexpr int MV_PAIR_FLOATS = 32;  // invK(9) R(9) t(3) Kj(9) pad(2)

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
                                    int T, int k_dim, int e_rows, float* __restrict__ pairs) {
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
}

__device__ __forceinline__ float fetch_zero_pad(const float* __restrict__ img, int x, int y, int W, int H) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(img + static_cast<long long>(y) * W + x) : 0.0f;
}

__global__ void __launch_bounds__(MV_THREADS)
mvcs_pairs_kernel(const float* __restrict__ depths, const float* __restrict__ pairs, int T, int H, int W,
                  int blocks_per_pair, double* __restrict__ part_sum, unsigned int* __restrict__ part_cnt) {
  const int pair_i = blockIdx.y;        // 0..T-2
  const int clip = blockIdx.z;
  const long long pair_idx = static_cast<long long>(clip) * (T - 1) + pair_i;
  // the 30 pair constants live in registers (the kernel is issue-bound: ~200 instructions per pixel, most of them the
  // reference's fp32 operation order without FMA contraction and four IEEE divisions)
  float sp[MV_PAIR_FLOATS - 2];
#pragma unroll
  for (int k = 0; k < MV_PAIR_FLOATS - 2; ++k) sp[k] = __ldg(pairs + pair_idx * MV_PAIR_FLOATS + k);
  const float* di = depths + (static_cast<long long>(clip) * T + pair_i) * H * W;
  const float* dj = di + static_cast<long long>(H) * W;
  const int HW = H * W;
  const float Wm1 = static_cast<float>(W - 1), Hm1 = static_cast<float>(H - 1);
  const float Wf = static_cast<float>(W), Hf = static_cast<float>(H);
  double acc = 0.0;
  unsigned int cnt = 0;
  const int stride = blocks_per_pair * MV_THREADS * MV_PIX_PER_THREAD;
  for (int base = (blockIdx.x * MV_THREADS + threadIdx.x) * MV_PIX_PER_THREAD; base < HW; base += stride) {
    float d4[MV_PIX_PER_THREAD];
    if (base + MV_PIX_PER_THREAD <= HW && (HW & 3) == 0) {
      const float4 q = *reinterpret_cast<const float4*>(di + base);
      d4[0] = q.x; d4[1] = q.y; d4[2] = q.z; d4[3] = q.w;
    } else {
#pragma unroll
      for (int k = 0; k < MV_PIX_PER_THREAD; ++k) d4[k] = (base + k < HW) ? di[base + k] : 0.f;
    }
    const bool one_row = (W & 3) == 0;                    // 4 consecutive pixels never straddle a row
    const int py0 = base / W, px0 = base - py0 * W;
#pragma unroll
    for (int k = 0; k < MV_PIX_PER_THREAD; ++k) {
      const int pix = base + k;
      if (pix >= HW) break;
      int py = py0, px = px0 + k;
      if (!one_row && px >= W) { py = pix / W; px = pix - py * W; }
      const float u = static_cast<float>(px), v = static_cast<float>(py), d = d4[k];
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
      if (in_mask) {
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
        const float tx = ix - fx, ty = iy - fy;
        const float w_nw = (1.0f - tx) * (1.0f - ty), w_ne = tx * (1.0f - ty);
        const float w_sw = (1.0f - tx) * ty, w_se = tx * ty;
        const float s = ((fetch_zero_pad(dj, x0, y0, W, H) * w_nw + fetch_zero_pad(dj, x0 + 1, y0, W, H) * w_ne) +
                         fetch_zero_pad(dj, x0, y0 + 1, W, H) * w_sw) + fetch_zero_pad(dj, x0 + 1, y0 + 1, W, H) * w_se;
        const float e = s - zj;
        acc += static_cast<double>(e * e);
        ++cnt;
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
  mvcs_prepare_kernel<<<(n_pairs + 127) / 128, 128, 0, s>>>(d_intrinsics, d_extrinsics, n_clips, T, k_dim, e_rows, pairs);
  VGPA_LAUNCH_CHECK("mvcs_prepare_kernel");
  VGPA_CHECK(n_clips <= 65535 && T - 1 <= 65535, "vgpa_mvcs_batch: too many clips per launch (%d)", n_clips);
  dim3 grid(bpp, T - 1, n_clips);
  mvcs_pairs_kernel<<<grid, MV_THREADS, 0, s>>>(d_depths, pairs, T, H, W, bpp, part_sum, part_cnt);
  VGPA_LAUNCH_CHECK("mvcs_pairs_kernel");
  mvcs_finalize_kernel<<<(n_clips + 127) / 128, 128, 0, s>>>(part_sum, part_cnt, n_clips, T, bpp, d_pair_mse,
                                                            reinterpret_cast<long long*>(d_pair_cnt), d_scores);
  VGPA_LAUNCH_CHECK("mvcs_finalize_kernel");
  return 0;
}
