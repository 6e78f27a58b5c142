// Consistency-score pieces that are pure geometry / reductions (the LPIPS network is an injected
// third-party callable):
//   - vgpa_motion_score      : compute_motion_score_vectorized (metrics/consistency_score.py:8-38)
//   - vgpa_mse_range_normalized : MSEMetric.compute with its value-dependent range heuristics
//                              (metrics/mse.py:14-54), one streaming pass over both videos
//   - vgpa_unproject_depth   : depth -> world points, DA3 path of VideoProcessor
//                              (pipelines/process_video.py:151-156; depth_anything_3/utils/geometry.py:54-59,434-498)
#include "common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {

// ------------------------------------------------------------------ motion score (T is ~10: one thread)
__global__ void motion_score_kernel(const float* __restrict__ E, int T, int e_rows, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int st = e_rows * 4;
  float sum_t = 0.f, sum_r = 0.f;
  for (int i = 0; i + 1 < T; ++i) {
    const float* a = E + static_cast<long long>(i) * st;       // E_i
    const float* b = E + static_cast<long long>(i + 1) * st;   // E_{i+1}
    const float dx = b[3] - a[3], dy = b[7] - a[7], dz = b[11] - a[11];
    sum_t += sqrtf((dx * dx + dy * dy) + dz * dz);                          // ||t_{i+1} - t_i||
    // trace(R_{i+1} R_i^T) = sum_rc R_{i+1}[r][c] * R_i[r][c], accumulated row by row like the batched matmul diagonal
    float tr = 0.f;
    for (int r = 0; r < 3; ++r) tr += (b[r * 4 + 0] * a[r * 4 + 0] + b[r * 4 + 1] * a[r * 4 + 1]) + b[r * 4 + 2] * a[r * 4 + 2];
    float c = (tr - 1.0f) / 2.0f;
    c = fminf(fmaxf(c, -1.0f), 1.0f);
    sum_r += acosf(c);
  }
  const float n = static_cast<float>(T - 1);
  const float score = sum_t / n + 0.1f * (sum_r / n);      // T == 1: 0/0 = NaN -> 0 below   (:36-37)
  out[0] = isnan(score) ? 0.0f : score;
}

// ------------------------------------------------------------------ range-normalised MSE
// moments per tensor pair: n, sum g, sum r, sum g^2, sum r^2, sum g r, plus min/max of each.
constexpr int MS_THREADS = 256;

struct MseSrc {
  const void* p;
  int kind;   // 0 = fp32, 1 = uint8
  int nhwc;   // layout of the source: 0 = [N, C, H, W], 1 = [N, H, W, C]
};

__device__ __forceinline__ float mse_fetch(const MseSrc& s, long long n, int c, long long pix, long long HW, int C) {
  const long long idx = s.nhwc ? (n * HW + pix) * C + c : (n * C + c) * HW + pix;
  return s.kind == 0 ? static_cast<const float*>(s.p)[idx] : static_cast<float>(static_cast<const unsigned char*>(s.p)[idx]);
}

__global__ void __launch_bounds__(MS_THREADS)
mse_moments_kernel(MseSrc g, MseSrc r, long long N, int C, long long HW, double* __restrict__ partial) {
  double sg = 0, sr = 0, sgg = 0, srr = 0, sgr = 0;
  float gmin = INFINITY, gmax = -INFINITY, rmin = INFINITY, rmax = -INFINITY;
  const long long total = N * C * HW;
  for (long long i = static_cast<long long>(blockIdx.x) * MS_THREADS + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * MS_THREADS) {
    // canonical order [N, C, H, W]
    const long long pix = i % HW;
    const int c = static_cast<int>((i / HW) % C);
    const long long n = i / (HW * C);
    const float a = mse_fetch(g, n, c, pix, HW, C), b = mse_fetch(r, n, c, pix, HW, C);
    sg += a; sr += b; sgg += static_cast<double>(a) * a; srr += static_cast<double>(b) * b; sgr += static_cast<double>(a) * b;
    gmin = fminf(gmin, a); gmax = fmaxf(gmax, a); rmin = fminf(rmin, b); rmax = fmaxf(rmax, b);
  }
  sg = warp_sum_d(sg); sr = warp_sum_d(sr); sgg = warp_sum_d(sgg); srr = warp_sum_d(srr); sgr = warp_sum_d(sgr);
  gmin = warp_min(gmin); gmax = warp_max(gmax); rmin = warp_min(rmin); rmax = warp_max(rmax);
  __shared__ double sh[MS_THREADS / 32][9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[warp][0] = sg; sh[warp][1] = sr; sh[warp][2] = sgg; sh[warp][3] = srr; sh[warp][4] = sgr;
    sh[warp][5] = gmin; sh[warp][6] = gmax; sh[warp][7] = rmin; sh[warp][8] = rmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double o[9];
    for (int k = 0; k < 9; ++k) o[k] = sh[0][k];
    for (int w = 1; w < MS_THREADS / 32; ++w) {
      for (int k = 0; k < 5; ++k) o[k] += sh[w][k];
      o[5] = fmin(o[5], sh[w][5]); o[6] = fmax(o[6], sh[w][6]); o[7] = fmin(o[7], sh[w][7]); o[8] = fmax(o[8], sh[w][8]);
    }
    for (int k = 0; k < 9; ++k) partial[static_cast<long long>(blockIdx.x) * 9 + k] = o[k];
  }
}

// t' = a t + b chosen like MSEMetric._to_tensor_01 (metrics/mse.py:31-54):
//   tensor input: min < 0 -> (t + 1) / 2 ; elif max > 1 -> t / 255 ; numpy input: max > 1 -> t / 255
__device__ void mse_affine(double mn, double mx, int is_numpy, double& a, double& b) {
  a = 1.0; b = 0.0;
  if (!is_numpy && mn < 0.0) { a = 0.5; b = 0.5; }
  else if (mx > 1.0) { a = 1.0 / 255.0; }
}

__global__ void mse_finalize_kernel(const double* __restrict__ partial, int n_blocks, double count, int g_numpy,
                                    int r_numpy, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double o[9];
  for (int k = 0; k < 9; ++k) o[k] = partial[k];
  for (int i = 1; i < n_blocks; ++i) {
    for (int k = 0; k < 5; ++k) o[k] += partial[static_cast<long long>(i) * 9 + k];
    o[5] = fmin(o[5], partial[static_cast<long long>(i) * 9 + 5]); o[6] = fmax(o[6], partial[static_cast<long long>(i) * 9 + 6]);
    o[7] = fmin(o[7], partial[static_cast<long long>(i) * 9 + 7]); o[8] = fmax(o[8], partial[static_cast<long long>(i) * 9 + 8]);
  }
  double ag, bg, ar, br;
  mse_affine(o[5], o[6], g_numpy, ag, bg);
  mse_affine(o[7], o[8], r_numpy, ar, br);
  // sum((ag g + bg - ar r - br)^2)
  const double d = bg - br;
  const double s = ag * ag * o[2] + ar * ar * o[3] - 2.0 * ag * ar * o[4] + 2.0 * d * (ag * o[0] - ar * o[1]) + count * d * d;
  out[0] = static_cast<float>(s / count);
}

// ------------------------------------------------------------------ depth -> world points
// world = c2w @ [ (K^-1 @ [x, y, 1]) * depth ; 1 ], c2w = affine_inverse(w2c) = [R^T | -R^T t]
__global__ void __launch_bounds__(256)
unproject_kernel(const float* __restrict__ depth, const float* __restrict__ Kmat, const float* __restrict__ Emat,
                 int T, int H, int W, int e_rows, float* __restrict__ out) {
  const int v = blockIdx.y;
  __shared__ float sK[9], sC[12];
  if (threadIdx.x == 0) {
    const float* K = Kmat + static_cast<long long>(v) * 9;
    // fp64 adjugate inverse rounded to fp32 (torch.inverse of a 3x3 in the reference)
    const double m[9] = {K[0], K[1], K[2], K[3], K[4], K[5], K[6], K[7], K[8]};
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double id = 1.0 / (m[0] * c00 + m[1] * c01 + m[2] * c02);
    sK[0] = static_cast<float>(c00 * id); sK[1] = static_cast<float>((m[2] * m[7] - m[1] * m[8]) * id);
    sK[2] = static_cast<float>((m[1] * m[5] - m[2] * m[4]) * id); sK[3] = static_cast<float>(c01 * id);
    sK[4] = static_cast<float>((m[0] * m[8] - m[2] * m[6]) * id); sK[5] = static_cast<float>((m[2] * m[3] - m[0] * m[5]) * id);
    sK[6] = static_cast<float>(c02 * id); sK[7] = static_cast<float>((m[1] * m[6] - m[0] * m[7]) * id);
    sK[8] = static_cast<float>((m[0] * m[4] - m[1] * m[3]) * id);
    const float* E = Emat + static_cast<long long>(v) * e_rows * 4;
    // affine_inverse: R^T and -R^T t   (depth_anything_3/utils/geometry.py:54-59)
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) sC[r * 4 + c] = E[c * 4 + r];
      sC[r * 4 + 3] = -((E[0 * 4 + r] * E[3] + E[1 * 4 + r] * E[7]) + E[2 * 4 + r] * E[11]);
    }
  }
  __syncthreads();
  const int hw = H * W;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += gridDim.x * blockDim.x) {
    const int py = pix / W, px = pix - py * W;
    const float x = static_cast<float>(px), y = static_cast<float>(py);
    const float d = depth[static_cast<long long>(v) * hw + pix];
    const float cx = ((sK[0] * x + sK[1] * y) + sK[2]) * d;
    const float cy = ((sK[3] * x + sK[4] * y) + sK[5]) * d;
    const float cz = ((sK[6] * x + sK[7] * y) + sK[8]) * d;
    float* o = out + (static_cast<long long>(v) * hw + pix) * 3;
    o[0] = ((sC[0] * cx + sC[1] * cy) + sC[2] * cz) + sC[3];
    o[1] = ((sC[4] * cx + sC[5] * cy) + sC[6] * cz) + sC[7];
    o[2] = ((sC[8] * cx + sC[9] * cy) + sC[10] * cz) + sC[11];
  }
}

}  // namespace
}  // namespace vgpa

extern "C" int vgpa_motion_score(const float* d_extrinsics, int T, int e_rows, float* d_out, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(d_extrinsics && d_out, "vgpa_motion_score: null pointer");
  VGPA_CHECK(T >= 1 && (e_rows == 3 || e_rows == 4), "vgpa_motion_score: bad shape T=%d e_rows=%d", T, e_rows);
  motion_score_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(d_extrinsics, T, e_rows, d_out);
  VGPA_LAUNCH_CHECK("motion_score_kernel");
  return 0;
}

extern "C" size_t vgpa_mse_workspace_bytes(void) { return static_cast<size_t>(148 * 8) * 9 * 8 + 256; }

extern "C" int vgpa_mse_range_normalized(const void* d_gt, int gt_kind, int gt_nhwc, int gt_numpy, const void* d_rep,
                                         int rep_kind, int rep_nhwc, int rep_numpy, int64_t N, int C, int H, int W,
                                         void* d_workspace, size_t workspace_bytes, float* d_out, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(d_gt && d_rep && d_out && d_workspace, "vgpa_mse_range_normalized: null pointer");
  VGPA_CHECK(N > 0 && C > 0 && H > 0 && W > 0, "vgpa_mse_range_normalized: bad shape");
  VGPA_CHECK((gt_kind == 0 || gt_kind == 1) && (rep_kind == 0 || rep_kind == 1), "vgpa_mse_range_normalized: kind must be 0 (fp32) or 1 (uint8)");
  VGPA_CHECK(workspace_bytes >= vgpa_mse_workspace_bytes(), "vgpa_mse_range_normalized: workspace too small");
  const long long total = static_cast<long long>(N) * C * H * W;
  long long nb = (total + MS_THREADS * 8 - 1) / (MS_THREADS * 8);
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb < 1) nb = 1;
  MseSrc g{d_gt, gt_kind, gt_nhwc}, r{d_rep, rep_kind, rep_nhwc};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(d_workspace);
  mse_moments_kernel<<<static_cast<unsigned>(nb), MS_THREADS, 0, s>>>(g, r, N, C, static_cast<long long>(H) * W, partial);
  VGPA_LAUNCH_CHECK("mse_moments_kernel");
  mse_finalize_kernel<<<1, 32, 0, s>>>(partial, static_cast<int>(nb), static_cast<double>(total), gt_numpy, rep_numpy, d_out);
  VGPA_LAUNCH_CHECK("mse_finalize_kernel");
  return 0;
}

extern "C" int vgpa_unproject_depth(const float* d_depth, const float* d_intrinsics, const float* d_extrinsics_w2c, int T,
                                    int H, int W, int e_rows, float* d_out_points, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(d_depth && d_intrinsics && d_extrinsics_w2c && d_out_points, "vgpa_unproject_depth: null pointer");
  VGPA_CHECK(T > 0 && T <= 65535 && H > 0 && W > 0 && (e_rows == 3 || e_rows == 4), "vgpa_unproject_depth: bad shape");
  int bx = (H * W + 255) / 256;
  if (bx > 1024) bx = 1024;
  unproject_kernel<<<dim3(bx, T), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_depth, d_intrinsics, d_extrinsics_w2c, T, H, W,
                                                                             e_rows, d_out_points);
  VGPA_LAUNCH_CHECK("unproject_kernel");
  return 0;
}
