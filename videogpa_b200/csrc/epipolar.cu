// K8: 8-point fundamental matrix + Sampson distance, one block per frame pair, all pairs in one launch.
//
// Replaces EpipolarMetric._compute_fundamental_matrix / _compute_sampson_distances
// (metrics/epipolar.py:194-216 of the reference), i.e. kornia.geometry.epipolar.find_fundamental
// (normalised 8-point: X^T X, eigenvector of the smallest eigenvalue, rank-2 projection,
// de-normalisation, F / (F22 + 1e-8)) and sampson_epipolar_distance(squared=True) followed by
// sqrt(d^2 + 1e-8) and the per-pair mean. The reference runs two batched GPU
// SVDs and a host sync per pair; here the 9x9 and 3x3 symmetric eigenproblems are solved with cyclic
// Jacobi in fp64 by one thread while the block does the streaming reductions around them. The
// keypoint matchers (SIFT / SuperPoint+LightGlue) stay third-party and feed this kernel.
#include "common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {

constexpr int EP_THREADS = 256;

__device__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < EP_THREADS / 32; ++w) t += sh[w];
  return t;
}

// cyclic Jacobi on a symmetric n x n matrix (row-major a, destroyed); v receives eigenvectors as columns
template <int N>
__device__ void jacobi_eig(double* a, double* v) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) v[i * N + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < N; ++i) {
      diag += a[i * N + i] * a[i * N + i];
      for (int j = i + 1; j < N; ++j) off += a[i * N + j] * a[i * N + j];
    }
    if (off <= 1e-60 || off <= 1e-32 * diag) break;
    for (int p = 0; p < N - 1; ++p) {
      for (int q = p + 1; q < N; ++q) {
        const double apq = a[p * N + q];
        if (apq == 0.0) continue;
        const double theta = (a[q * N + q] - a[p * N + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {
          const double akp = a[k * N + p], akq = a[k * N + q];
          a[k * N + p] = c * akp - s * akq;
          a[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {
          const double apk = a[p * N + k], aqk = a[q * N + k];
          a[p * N + k] = c * apk - s * aqk;
          a[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = v[k * N + p], vkq = v[k * N + q];
          v[k * N + p] = c * vkp - s * vkq;
          v[k * N + q] = s * vkp + c * vkq;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(EP_THREADS)
epipolar_kernel(const float* __restrict__ pts1, const float* __restrict__ pts2, const int* __restrict__ counts,
                int max_matches, float* __restrict__ F_out, float* __restrict__ dist_out, int* __restrict__ valid_out) {
  const int pair = blockIdx.x;
  const int n = counts ? counts[pair] : max_matches;
  const float* p1 = pts1 + static_cast<long long>(pair) * max_matches * 2;
  const float* p2 = pts2 + static_cast<long long>(pair) * max_matches * 2;
  __shared__ double sh[EP_THREADS / 32];
  __shared__ double sA[81];
  __shared__ double sT[8];     // s1, mx1, my1, s2, mx2, my2
  __shared__ float sF[9];
  __shared__ int s_ok;
  if (n < 8 || n > max_matches) {
    if (threadIdx.x == 0) {
      valid_out[pair] = 0;
      dist_out[pair] = nanf("");
      for (int k = 0; k < 9; ++k) F_out[pair * 9 + k] = nanf("");
    }
    return;
  }
  // ---- normalisation transforms: centroid, scale = sqrt(2) / (mean distance + 1e-8)
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int i = threadIdx.x; i < n; i += EP_THREADS) { a0 += p1[2 * i]; a1 += p1[2 * i + 1]; a2 += p2[2 * i]; a3 += p2[2 * i + 1]; }
  const double mx1 = block_sum_d(a0, sh) / n, my1 = block_sum_d(a1, sh) / n;
  const double mx2 = block_sum_d(a2, sh) / n, my2 = block_sum_d(a3, sh) / n;
  double d1 = 0, d2 = 0;
  for (int i = threadIdx.x; i < n; i += EP_THREADS) {
    const double ax = p1[2 * i] - mx1, ay = p1[2 * i + 1] - my1, bx = p2[2 * i] - mx2, by = p2[2 * i + 1] - my2;
    d1 += sqrt(ax * ax + ay * ay);
    d2 += sqrt(bx * bx + by * by);
  }
  const double s1 = 1.4142135623730951 / (block_sum_d(d1, sh) / n + 1e-8);
  const double s2 = 1.4142135623730951 / (block_sum_d(d2, sh) / n + 1e-8);
  // ---- A = X^T X with X_i = [x2x1, x2y1, x2, y2x1, y2y1, y2, x1, y1, 1]
  double acc[45];
#pragma unroll
  for (int k = 0; k < 45; ++k) acc[k] = 0.0;
  for (int i = threadIdx.x; i < n; i += EP_THREADS) {
    const double x1 = (p1[2 * i] - mx1) * s1, y1 = (p1[2 * i + 1] - my1) * s1;
    const double x2 = (p2[2 * i] - mx2) * s2, y2 = (p2[2 * i + 1] - my2) * s2;
    const double X[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
    int k = 0;
#pragma unroll
    for (int r = 0; r < 9; ++r)
#pragma unroll
      for (int c = r; c < 9; ++c) acc[k++] += X[r] * X[c];
  }
  {
    int k = 0;
    for (int r = 0; r < 9; ++r)
      for (int c = r; c < 9; ++c) {
        const double t = block_sum_d(acc[k++], sh);
        if (threadIdx.x == 0) { sA[r * 9 + c] = t; sA[c * 9 + r] = t; }
      }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double V[81];
    jacobi_eig<9>(sA, V);
    int kmin = 0;
    for (int k = 1; k < 9; ++k) if (sA[k * 9 + k] < sA[kmin * 9 + kmin]) kmin = k;
    double Fh[9];
    for (int k = 0; k < 9; ++k) Fh[k] = V[k * 9 + kmin];
    // rank-2 projection: drop the smallest singular value, F' = F - (F v3) v3^T with v3 from F^T F
    double G[9], W3[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) G[r * 3 + c] = Fh[0 * 3 + r] * Fh[0 * 3 + c] + Fh[1 * 3 + r] * Fh[1 * 3 + c] + Fh[2 * 3 + r] * Fh[2 * 3 + c];
    jacobi_eig<3>(G, W3);
    int k3 = 0;
    for (int k = 1; k < 3; ++k) if (G[k * 3 + k] < G[k3 * 3 + k3]) k3 = k;
    const double v3[3] = {W3[0 * 3 + k3], W3[1 * 3 + k3], W3[2 * 3 + k3]};
    double Fp[9];
    for (int r = 0; r < 3; ++r) {
      const double fv = Fh[r * 3 + 0] * v3[0] + Fh[r * 3 + 1] * v3[1] + Fh[r * 3 + 2] * v3[2];
      for (int c = 0; c < 3; ++c) Fp[r * 3 + c] = Fh[r * 3 + c] - fv * v3[c];
    }
    // F = T2^T F' T1, T = [[s, 0, -s mx], [0, s, -s my], [0, 0, 1]]
    const double T1[9] = {s1, 0, -s1 * mx1, 0, s1, -s1 * my1, 0, 0, 1};
    const double T2[9] = {s2, 0, -s2 * mx2, 0, s2, -s2 * my2, 0, 0, 1};
    double M[9], Fd[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) M[r * 3 + c] = Fp[r * 3 + 0] * T1[0 * 3 + c] + Fp[r * 3 + 1] * T1[1 * 3 + c] + Fp[r * 3 + 2] * T1[2 * 3 + c];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Fd[r * 3 + c] = T2[0 * 3 + r] * M[0 * 3 + c] + T2[1 * 3 + r] * M[1 * 3 + c] + T2[2 * 3 + r] * M[2 * 3 + c];
    // normalize_transformation: M / (M22 + eps) where |M22| > eps
    const double f22 = Fd[8];
    int ok = 1;
    for (int k = 0; k < 9; ++k) {
      const double f = (fabs(f22) > 1e-8) ? Fd[k] / (f22 + 1e-8) : Fd[k];
      sF[k] = static_cast<float>(f);
      if (isnan(sF[k])) ok = 0;                 // NaN -> pair skipped (metrics/epipolar.py:202)
      F_out[pair * 9 + k] = sF[k];
    }
    s_ok = ok;
  }
  __syncthreads();
  // ---- Sampson distances in fp32 (as kornia does on the fp32 points), mean of sqrt(d^2 + 1e-8)
  double dsum = 0.0;
  for (int i = threadIdx.x; i < n; i += EP_THREADS) {
    const float x1 = p1[2 * i], y1 = p1[2 * i + 1], x2 = p2[2 * i], y2 = p2[2 * i + 1];
    const float l0 = (sF[0] * x1 + sF[1] * y1) + sF[2];        // F p1
    const float l1 = (sF[3] * x1 + sF[4] * y1) + sF[5];
    const float l2 = (sF[6] * x1 + sF[7] * y1) + sF[8];
    const float m0 = (sF[0] * x2 + sF[3] * y2) + sF[6];        // F^T p2
    const float m1 = (sF[1] * x2 + sF[4] * y2) + sF[7];
    const float num = (x2 * l0 + y2 * l1) + l2;
    const float den = ((l0 * l0 + l1 * l1) + m0 * m0) + m1 * m1;
    const float d2v = (num * num) / den;
    dsum += static_cast<double>(sqrtf(d2v + 1e-8f));            // metrics/epipolar.py:213
  }
  const double tot = block_sum_d(dsum, sh);
  if (threadIdx.x == 0) {
    valid_out[pair] = s_ok;
    dist_out[pair] = s_ok ? static_cast<float>(tot / n) : nanf("");
  }
}

}  // namespace
}  // namespace vgpa

extern "C" int vgpa_epipolar_batch(const float* d_pts1, const float* d_pts2, const int32_t* d_counts, int n_pairs,
                                   int max_matches, float* d_F, float* d_mean_dist, int32_t* d_valid, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(n_pairs >= 0 && max_matches > 0, "vgpa_epipolar_batch: bad shape pairs=%d max_matches=%d", n_pairs, max_matches);
  if (n_pairs == 0) return 0;
  VGPA_CHECK(d_pts1 && d_pts2 && d_F && d_mean_dist && d_valid, "vgpa_epipolar_batch: null pointer");
  epipolar_kernel<<<n_pairs, EP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(d_pts1, d_pts2, d_counts, max_matches, d_F,
                                                                                 d_mean_dist, d_valid);
  VGPA_LAUNCH_CHECK("epipolar_kernel");
  return 0;
}
