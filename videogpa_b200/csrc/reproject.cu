// K6: point-cloud reprojection renderer — every view of a clip in three launches.
//
// Replaces project_points / batch_reproject (utils/projection_utils.py:12-101 of the reference; called
// from pipelines/process_video.py:85,117). The reference renders each view with two matmuls, an
// argsort of all N depths (painter's algorithm) and a last-write-wins scatter. Here nearest-z wins
// through a 64-bit z-buffer: key = (float_bits(z) << 32) | point_index, resolved with atomicMin, so
//   - the nearest point wins, and on an exact z tie the lowest point index wins (deterministic,
//     where the reference's unstable argsort + duplicate index_put is not),
//   - each point is read once for all T views (12 B/point), not once per view.
// Pixel indices are integer work and must be bit-exact against the oracle: the projection keeps the
// reference's operation order in fp32 and this file is compiled with --fmad=false.
#include "common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {

constexpr int RP_MAX_VIEWS = 32;
constexpr unsigned long long RP_EMPTY = 0xFFFFFFFFFFFFFFFFull;

struct RpViews {
  float R[RP_MAX_VIEWS][9];
  float t[RP_MAX_VIEWS][3];
  float K[RP_MAX_VIEWS][9];
};

// order-preserving float <-> uint mapping so that atomicMax works on any sign
__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

__global__ void rp_init_kernel(unsigned long long* zbuf, long long n, unsigned int* cmax, int T) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) zbuf[i] = RP_EMPTY;
  if (i < T) cmax[i] = 0u;  // below every encoded float
}

__global__ void __launch_bounds__(256)
rp_project_kernel(const float* __restrict__ pc, const float* __restrict__ colors, const float* __restrict__ Kmat,
                  const float* __restrict__ Emat, long long P, int T, int H, int W, int e_rows,
                  unsigned long long* __restrict__ zbuf, unsigned int* __restrict__ cmax) {
  __shared__ float sR[RP_MAX_VIEWS][9];
  __shared__ float sT[RP_MAX_VIEWS][3];
  __shared__ float sK[RP_MAX_VIEWS][9];
  __shared__ unsigned int s_cmax[RP_MAX_VIEWS];
  for (int i = threadIdx.x; i < T * 9; i += blockDim.x) {
    const int v = i / 9, k = i - v * 9;
    sR[v][k] = Emat[static_cast<long long>(v) * e_rows * 4 + (k / 3) * 4 + (k % 3)];
    sK[v][k] = Kmat[static_cast<long long>(v) * 9 + k];
  }
  for (int i = threadIdx.x; i < T * 3; i += blockDim.x) {
    const int v = i / 3, k = i - v * 3;
    sT[v][k] = Emat[static_cast<long long>(v) * e_rows * 4 + k * 4 + 3];
  }
  if (threadIdx.x < RP_MAX_VIEWS) s_cmax[threadIdx.x] = 0u;
  __syncthreads();
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < P) {
    const float x = pc[p * 3 + 0], y = pc[p * 3 + 1], z = pc[p * 3 + 2];
    const float c0 = colors[p * 3 + 0], c1 = colors[p * 3 + 1], c2 = colors[p * 3 + 2];
    const unsigned int cm = f2ord(fmaxf(fmaxf(c0, c1), c2));
    for (int v = 0; v < T; ++v) {
      // pc_cam = pc @ R^T + t ; pc_proj = pc_cam @ K^T        (projection_utils.py:19-20)
      const float xc = ((x * sR[v][0] + y * sR[v][1]) + z * sR[v][2]) + sT[v][0];
      const float yc = ((x * sR[v][3] + y * sR[v][4]) + z * sR[v][5]) + sT[v][1];
      const float zc = ((x * sR[v][6] + y * sR[v][7]) + z * sR[v][8]) + sT[v][2];
      const float px = (xc * sK[v][0] + yc * sK[v][1]) + zc * sK[v][2];
      const float py = (xc * sK[v][3] + yc * sK[v][4]) + zc * sK[v][5];
      const float pz = (xc * sK[v][6] + yc * sK[v][7]) + zc * sK[v][8];
      // u, v = round-half-even(x / (z + 1e-8))                (projection_utils.py:22-24)
      const float den = pz + 1e-8f;
      const float uf = rintf(px / den), vf = rintf(py / den);
      // valid = in bounds and z > 0                            (projection_utils.py:26)
      if (uf >= 0.0f && uf < static_cast<float>(W) && vf >= 0.0f && vf < static_cast<float>(H) && pz > 0.0f) {
        const int ui = static_cast<int>(uf), vi = static_cast<int>(vf);
        const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(pz)) << 32) |
                                       static_cast<unsigned long long>(static_cast<unsigned int>(p));
        atomicMin(&zbuf[(static_cast<long long>(v) * H + vi) * W + ui], key);
        atomicMax(&s_cmax[v], cm);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < T && s_cmax[threadIdx.x] != 0u) atomicMax(&cmax[threadIdx.x], s_cmax[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
rp_resolve_kernel(const unsigned long long* __restrict__ zbuf, const float* __restrict__ colors,
                  const unsigned int* __restrict__ cmax, int T, int H, int W, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long hw = static_cast<long long>(H) * W;
  if (i >= hw * T) return;
  const int v = static_cast<int>(i / hw);
  const long long pix = i - v * hw;
  const unsigned long long key = zbuf[i];
  float r[3] = {0.f, 0.f, 0.f};  // background (0, 0, 0)                (projection_utils.py:43)
  if (key != RP_EMPTY) {
    const unsigned int idx = static_cast<unsigned int>(key & 0xFFFFFFFFull);
    // colours in [0, 1] are scaled by 255, decided by the max over the view's valid points (:45-48)
    const bool unit = ord2f(cmax[v]) <= 1.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float f = colors[static_cast<long long>(idx) * 3 + c];
      if (unit) f = f * 255.0f;
      f = fminf(fmaxf(f, 0.0f), 255.0f);                 // NaN -> 0 through fmaxf
      r[c] = static_cast<float>(static_cast<unsigned char>(f));  // truncation, like .to(torch.uint8)
    }
  }
  // stack -> [T, 3, H, W] float, (x / 255) * 2 - 1                      (projection_utils.py:100-101)
#pragma unroll
  for (int c = 0; c < 3; ++c) out[(static_cast<long long>(v) * 3 + c) * hw + pix] = (r[c] / 255.0f) * 2.0f - 1.0f;
}

}  // namespace
}  // namespace vgpa

extern "C" size_t vgpa_reproject_workspace_bytes(int T, int H, int W) {
  return static_cast<size_t>(T) * H * W * 8 + 256;
}

extern "C" int vgpa_reproject_batch(const float* d_points, const float* d_colors, const float* d_intrinsics,
                                    const float* d_extrinsics, int64_t n_points, int T, int H, int W, int e_rows,
                                    void* d_workspace, size_t workspace_bytes, float* d_out, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(T >= 0 && T <= RP_MAX_VIEWS, "vgpa_reproject_batch: T=%d must be in [0, %d]", T, RP_MAX_VIEWS);
  VGPA_CHECK(H > 0 && W > 0 && n_points >= 0, "vgpa_reproject_batch: bad shape H=%d W=%d N=%lld", H, W, (long long)n_points);
  VGPA_CHECK(n_points <= 0xFFFFFFFFll, "vgpa_reproject_batch: at most 2^32-1 points per call");
  VGPA_CHECK(e_rows == 3 || e_rows == 4, "vgpa_reproject_batch: extrinsics must be 3x4 or 4x4");
  if (T == 0) return 0;
  VGPA_CHECK(d_intrinsics && d_extrinsics && d_out && d_workspace, "vgpa_reproject_batch: null pointer");
  VGPA_CHECK(n_points == 0 || (d_points && d_colors), "vgpa_reproject_batch: null point/colour pointer");
  VGPA_CHECK(workspace_bytes >= vgpa_reproject_workspace_bytes(T, H, W), "vgpa_reproject_batch: workspace too small");
  VGPA_CHECK((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, "vgpa_reproject_batch: workspace must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long npix = static_cast<long long>(T) * H * W;
  unsigned long long* zbuf = static_cast<unsigned long long*>(d_workspace);
  unsigned int* cmax = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(d_workspace) + npix * 8);
  rp_init_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, s>>>(zbuf, npix, cmax, T);
  VGPA_LAUNCH_CHECK("rp_init_kernel");
  if (n_points > 0) {
    rp_project_kernel<<<static_cast<unsigned>((n_points + 255) / 256), 256, 0, s>>>(
        d_points, d_colors, d_intrinsics, d_extrinsics, n_points, T, H, W, e_rows, zbuf, cmax);
    VGPA_LAUNCH_CHECK("rp_project_kernel");
  }
  rp_resolve_kernel<<<static_cast<unsigned>((npix + 255) / 256), 256, 0, s>>>(zbuf, d_colors, cmax, T, H, W, d_out);
  VGPA_LAUNCH_CHECK("rp_resolve_kernel");
  return 0;
}
