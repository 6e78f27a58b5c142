// K7: coloured point cloud assembly — confidence filter (finite, > 1e-5, optional top-percentile) and
// order-preserving stream compaction of (vertex, colour) pairs.
//
// Replaces get_colored_pointcloud (utils/pointcloud_utils.py:10-80 of the reference; called from
// pipelines/process_video.py:84,116). The reference runs torch.topk over millions of confidences plus a
// `.item()` sync to find the k-th value; here the threshold is an exact 4-pass (8 bits each) radix
// select on the float bit patterns (valid confidences are positive, so the bits order like the
// values), entirely on the device. Ties at the threshold are kept (`vals >= thr`, :73), and the
// survivors keep their original order, exactly like boolean-mask indexing.
// Colours are gathered from images [T, 3, H, W] in [0, 1] as (image.permute(0,2,3,1) * 255) (:36-42).
#include "common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {

constexpr int PC_THREADS = 256;
constexpr int PC_ITEMS = 4;                       // elements per thread
constexpr int PC_TILE = PC_THREADS * PC_ITEMS;    // 1024 elements per block

struct PcState {               // device-resident control block
  unsigned long long n_valid;
  unsigned long long k_remaining;
  unsigned int prefix;
  unsigned int prefix_mask;
  unsigned int threshold_bits;
  unsigned int pad;
  unsigned long long n_kept;
  unsigned int hist[256];
};

__device__ __forceinline__ bool pc_valid(float v) { return isfinite(v) && v > 1e-5f; }   // pointcloud_utils.py:47

__global__ void pc_reset_kernel(PcState* st) {
  if (threadIdx.x == 0) {
    st->n_valid = 0; st->k_remaining = 0; st->prefix = 0; st->prefix_mask = 0; st->threshold_bits = 0; st->n_kept = 0;
  }
  st->hist[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(PC_THREADS)
pc_count_valid_kernel(const float* __restrict__ conf, long long n, PcState* st) {
  unsigned int c = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    c += pc_valid(conf[i]) ? 1u : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&st->n_valid, static_cast<unsigned long long>(c));
}

// k = max(1, ceil(N_valid * keep_frac))                                 (pointcloud_utils.py:60-61)
__global__ void pc_set_k_kernel(PcState* st, double keep_frac) {
  const double nv = static_cast<double>(st->n_valid);
  long long k = static_cast<long long>(ceil(nv * keep_frac));
  if (k < 1) k = 1;
  st->k_remaining = static_cast<unsigned long long>(k);
}

__global__ void __launch_bounds__(PC_THREADS)
pc_hist_kernel(const float* __restrict__ conf, long long n, PcState* st, int shift) {
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const unsigned int prefix = st->prefix, mask = st->prefix_mask;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = conf[i];
    if (pc_valid(v)) {
      const unsigned int b = __float_as_uint(v);
      if ((b & mask) == prefix) atomicAdd(&sh[(b >> shift) & 0xFFu], 1u);
    }
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], sh[threadIdx.x]);
}

// walk the histogram from the largest digit down until the k-th largest value's digit is found
__global__ void pc_select_kernel(PcState* st, int shift) {
  if (threadIdx.x == 0) {
    unsigned long long k = st->k_remaining;
    int d = 255;
    for (; d > 0; --d) {
      const unsigned long long h = st->hist[d];
      if (k <= h) break;
      k -= h;
    }
    st->k_remaining = k;
    st->prefix |= static_cast<unsigned int>(d) << shift;
    st->prefix_mask |= 0xFFu << shift;
    if (shift == 0) st->threshold_bits = st->prefix;
  }
  __syncthreads();
  st->hist[threadIdx.x] = 0;
}

__device__ __forceinline__ bool pc_keep(float v, bool use_thr, float thr) {
  return pc_valid(v) && (!use_thr || v >= thr);
}

__global__ void __launch_bounds__(PC_THREADS)
pc_block_count_kernel(const float* __restrict__ conf, long long n, const PcState* __restrict__ st, int use_thr,
                      unsigned int* __restrict__ block_counts) {
  const float thr = __uint_as_float(st->threshold_bits);
  const bool skip = use_thr && st->n_valid == 0;   // N == 0: mask stays all-false (pointcloud_utils.py:56-58)
  const long long base = static_cast<long long>(blockIdx.x) * PC_TILE + threadIdx.x * PC_ITEMS;
  unsigned int c = 0;
#pragma unroll
  for (int k = 0; k < PC_ITEMS; ++k)
    if (base + k < n && !skip && pc_keep(conf[base + k], use_thr, thr)) ++c;
  __shared__ unsigned int ws[PC_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (int w = 0; w < PC_THREADS / 32; ++w) t += ws[w];
    block_counts[blockIdx.x] = t;
  }
}

// single-block exclusive scan of the per-block counts (n_blocks is a few thousand)
__global__ void __launch_bounds__(1024)
pc_scan_kernel(const unsigned int* __restrict__ block_counts, unsigned long long* __restrict__ block_offsets,
               int n_blocks, PcState* st, long long* __restrict__ out_count) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned long long v = (i < n_blocks) ? block_counts[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long w = warp_tot[threadIdx.x];
      unsigned long long wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += t;
      }
      warp_tot[threadIdx.x] = wi - w;  // exclusive
    }
    __syncthreads();
    const unsigned long long excl = carry + warp_tot[threadIdx.x >> 5] + (incl - v);
    if (i < n_blocks) block_offsets[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) { st->n_kept = carry; *out_count = static_cast<long long>(carry); }
}

__global__ void __launch_bounds__(PC_THREADS)
pc_scatter_kernel(const float* __restrict__ points, const float* __restrict__ images, const float* __restrict__ conf,
                  long long n, int HW, int nhwc, const PcState* __restrict__ st, int use_thr,
                  const unsigned long long* __restrict__ block_offsets, float* __restrict__ out_v,
                  float* __restrict__ out_c) {
  const float thr = __uint_as_float(st->threshold_bits);
  const bool skip = use_thr && st->n_valid == 0;
  const long long base = static_cast<long long>(blockIdx.x) * PC_TILE + threadIdx.x * PC_ITEMS;
  bool keep[PC_ITEMS];
  unsigned int c = 0;
#pragma unroll
  for (int k = 0; k < PC_ITEMS; ++k) {
    keep[k] = (base + k < n) && !skip && pc_keep(conf[base + k], use_thr, thr);
    c += keep[k] ? 1u : 0u;
  }
  // exclusive scan of per-thread counts across the block
  unsigned int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  __shared__ unsigned int ws[PC_THREADS / 32];
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
  __syncthreads();
  unsigned int woff = 0;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += ws[w];
  unsigned long long dst = block_offsets[blockIdx.x] + woff + (incl - c);
#pragma unroll
  for (int k = 0; k < PC_ITEMS; ++k) {
    if (keep[k]) {
      const long long i = base + k;
      out_v[dst * 3 + 0] = points[i * 3 + 0];
      out_v[dst * 3 + 1] = points[i * 3 + 1];
      out_v[dst * 3 + 2] = points[i * 3 + 2];
      // colours = images[T,3,H,W].permute(0,2,3,1).reshape(-1,3) * 255
      const long long t = i / HW, pix = i - t * HW;
      if (nhwc) {
        out_c[dst * 3 + 0] = images[i * 3 + 0] * 255.0f;
        out_c[dst * 3 + 1] = images[i * 3 + 1] * 255.0f;
        out_c[dst * 3 + 2] = images[i * 3 + 2] * 255.0f;
      } else {
        out_c[dst * 3 + 0] = images[(t * 3 + 0) * HW + pix] * 255.0f;
        out_c[dst * 3 + 1] = images[(t * 3 + 1) * HW + pix] * 255.0f;
        out_c[dst * 3 + 2] = images[(t * 3 + 2) * HW + pix] * 255.0f;
      }
      ++dst;
    }
  }
}

}  // namespace
}  // namespace vgpa

extern "C" size_t vgpa_pointcloud_workspace_bytes(int64_t n_points) {
  const long long n_blocks = (n_points + vgpa::PC_TILE - 1) / vgpa::PC_TILE;
  return 2048 + static_cast<size_t>(n_blocks > 0 ? n_blocks : 1) * 16 + 256;
}

extern "C" int vgpa_pointcloud_filter(const float* d_points, const float* d_images, const float* d_conf,
                                      int64_t n_points, int hw_per_frame, int images_nhwc, double conf_thres, void* d_workspace,
                                      size_t workspace_bytes, float* d_out_vertices, float* d_out_colors,
                                      int64_t* d_out_count, float* d_out_threshold, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(n_points >= 0 && hw_per_frame > 0, "vgpa_pointcloud_filter: bad shape n=%lld hw=%d", (long long)n_points, hw_per_frame);
  VGPA_CHECK(d_workspace && d_out_count, "vgpa_pointcloud_filter: null workspace / count pointer");
  VGPA_CHECK(workspace_bytes >= vgpa_pointcloud_workspace_bytes(n_points), "vgpa_pointcloud_filter: workspace too small");
  VGPA_CHECK((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, "vgpa_pointcloud_filter: workspace must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_points == 0) {
    VGPA_CUDA(cudaMemsetAsync(d_out_count, 0, sizeof(int64_t), s));
    return 0;
  }
  VGPA_CHECK(d_points && d_images && d_conf && d_out_vertices && d_out_colors, "vgpa_pointcloud_filter: null tensor pointer");
  PcState* st = static_cast<PcState*>(d_workspace);
  const int n_blocks = static_cast<int>((n_points + PC_TILE - 1) / PC_TILE);
  unsigned int* block_counts = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(d_workspace) + 2048);
  unsigned long long* block_offsets = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(d_workspace) + 2048 +
                                                                            ((static_cast<size_t>(n_blocks) * 4 + 7) & ~size_t(7)));
  const int use_thr = conf_thres > 0.0 ? 1 : 0;
  pc_reset_kernel<<<1, 256, 0, s>>>(st);
  VGPA_LAUNCH_CHECK("pc_reset_kernel");
  if (use_thr) {
    int sweep = num_sms() * 8;
    if (sweep > n_blocks * PC_ITEMS) sweep = n_blocks * PC_ITEMS;
    pc_count_valid_kernel<<<sweep, PC_THREADS, 0, s>>>(d_conf, n_points, st);
    VGPA_LAUNCH_CHECK("pc_count_valid_kernel");
    double keep_frac = 1.0 - conf_thres / 100.0;                       // pointcloud_utils.py:60
    keep_frac = keep_frac < 0.0 ? 0.0 : (keep_frac > 1.0 ? 1.0 : keep_frac);
    pc_set_k_kernel<<<1, 1, 0, s>>>(st, keep_frac);
    VGPA_LAUNCH_CHECK("pc_set_k_kernel");
    for (int shift = 24; shift >= 0; shift -= 8) {
      pc_hist_kernel<<<sweep, PC_THREADS, 0, s>>>(d_conf, n_points, st, shift);
      VGPA_LAUNCH_CHECK("pc_hist_kernel");
      pc_select_kernel<<<1, 256, 0, s>>>(st, shift);
      VGPA_LAUNCH_CHECK("pc_select_kernel");
    }
  }
  pc_block_count_kernel<<<n_blocks, PC_THREADS, 0, s>>>(d_conf, n_points, st, use_thr, block_counts);
  VGPA_LAUNCH_CHECK("pc_block_count_kernel");
  pc_scan_kernel<<<1, 1024, 0, s>>>(block_counts, block_offsets, n_blocks, st, reinterpret_cast<long long*>(d_out_count));
  VGPA_LAUNCH_CHECK("pc_scan_kernel");
  pc_scatter_kernel<<<n_blocks, PC_THREADS, 0, s>>>(d_points, d_images, d_conf, n_points, hw_per_frame, images_nhwc, st, use_thr,
                                                    block_offsets, d_out_vertices, d_out_colors);
  VGPA_LAUNCH_CHECK("pc_scatter_kernel");
  if (d_out_threshold) {
    VGPA_CUDA(cudaMemcpyAsync(d_out_threshold, &st->threshold_bits, 4, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}
