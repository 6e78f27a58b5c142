// K4b: the HBM-bound half of the CogVideoX VAE decoder — GroupNorm statistics, SpatialNorm3D apply (+SiLU),
// nearest-neighbour upsampling and the tiled-decode blend/crop.
//
// Replaces torch.nn.GroupNorm / F.interpolate / F.silu inside diffusers' CogVideoXSpatialNorm3D,
// CogVideoXResnetBlock3D, CogVideoXUpsample3D and AutoencoderKLCogVideoX.tiled_decode (SURVEY.md App. A.5;
// reference call site generate/CogVideoX-5B.py:20-21,72-77). Activations are channels-last [T, H, W, C] bf16:
// every kernel moves 16-byte (8-channel) vectors, a warp covering 256 contiguous channels-bytes of one pixel.
#include "common.cuh"
#include "../../include/videogpa_b200.h"

namespace vgpa {
namespace {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_BLOCKS = 592;   // 148 SMs x 4

// ---------------------------------------------------------------------------------------------- GroupNorm stats
// Pass 1: every block accumulates per-channel sum / sum of squares over a strided set of pixels (fp32 per thread,
// combined across the block in shared memory) and writes one [2, C] partial. Pass 2 (one block) folds the
// partials in fp64 in a fixed order (deterministic) and reduces channels to groups.
__global__ void __launch_bounds__(GN_THREADS)
gn_partial_kernel(const __nv_bfloat16* __restrict__ x, long long n_pixels, int C, float* __restrict__ partial) {
  extern __shared__ float sh[];                    // [2][C]
  const int vec_per_pixel = C / 8;                 // uint4 vectors per pixel
  const int pix_per_iter = GN_THREADS / vec_per_pixel;
  const int v = threadIdx.x % vec_per_pixel;       // this thread's 8 channels
  const int pslot = threadIdx.x / vec_per_pixel;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (pslot < pix_per_iter) {
    const long long stride = static_cast<long long>(gridDim.x) * pix_per_iter;
    const __nv_bfloat16* xv = x + v * 8;
    auto fold = [&](const uint4& u) {
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16x2(w[i]);
        s[2 * i] += f.x; q[2 * i] += f.x * f.x;
        s[2 * i + 1] += f.y; q[2 * i + 1] += f.y * f.y;
      }
    };
    long long pix = static_cast<long long>(blockIdx.x) * pix_per_iter + pslot;
    // four independent 16-byte loads in flight per thread; the fold order (increasing pixel) is unchanged
    for (; pix + 3 * stride < n_pixels; pix += 4 * stride) {
      const uint4 u0 = *reinterpret_cast<const uint4*>(xv + pix * C);
      const uint4 u1 = *reinterpret_cast<const uint4*>(xv + (pix + stride) * C);
      const uint4 u2 = *reinterpret_cast<const uint4*>(xv + (pix + 2 * stride) * C);
      const uint4 u3 = *reinterpret_cast<const uint4*>(xv + (pix + 3 * stride) * C);
      fold(u0); fold(u1); fold(u2); fold(u3);
    }
    for (; pix < n_pixels; pix += stride) fold(*reinterpret_cast<const uint4*>(xv + pix * C));
  }
  for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) sh[i] = 0.f;
  __syncthreads();
  // pixel slots are folded one after another so the order of additions is fixed
  for (int turn = 0; turn < pix_per_iter; ++turn) {
    if (pslot == turn) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { sh[v * 8 + i] += s[i]; sh[C + v * 8 + i] += q[i]; }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) partial[static_cast<long long>(blockIdx.x) * 2 * C + i] = sh[i];
}

// One block per group: thread i folds partial blocks i, i+256, ... in fp64, then a fixed shared-memory tree.
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, int groups, double count,
                   float eps, float* __restrict__ mean_rstd) {
  __shared__ double ss[256], sq[256];
  const int g = blockIdx.x;
  const int cpg = C / groups;
  double s = 0.0, q = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 256) {
    const float* p = partial + static_cast<long long>(b) * 2 * C + g * cpg;
    for (int c = 0; c < cpg; ++c) { s += p[c]; q += p[C + c]; }
  }
  ss[threadIdx.x] = s;
  sq[threadIdx.x] = q;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) { ss[threadIdx.x] += ss[threadIdx.x + off]; sq[threadIdx.x] += sq[threadIdx.x + off]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = ss[0] / count;
    double var = sq[0] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_rstd[g] = static_cast<float>(mean);
    mean_rstd[groups + g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

// ---------------------------------------------------------------------------------------------- SpatialNorm apply
struct SnParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  const float* mean_rstd;
  const __nv_bfloat16* gamma;
  const __nv_bfloat16* beta;
  const __nv_bfloat16* y_lat;
  const __nv_bfloat16* b_lat;
  long long ld_lat;
  int T, H, W, C, groups, Hz, Wz, shift, silu;
  int tz_of_t[16];
};

// Thread mapping shared by the apply / upsample kernels: a block walks rows (t, h) of the [T, H, W, C] tensor; thread
// `tid` owns the 8 channels v = tid % (C/8) of pixels slot, slot + 256/(C/8), ... of the row, so everything that depends
// on the channel (GroupNorm scale/shift) is loop-invariant and no per-element integer division is left in the loop.
// blockIdx.y splits long rows when there are too few of them to fill the SMs.
__global__ void __launch_bounds__(256, 4)
spatialnorm_apply_kernel(SnParams p) {
  const int C = p.C, vpp = C >> 3, ppi = 256 / vpp;
  const int v = threadIdx.x % vpp, slot = threadIdx.x / vpp;
  if (slot >= ppi) return;
  const int c0 = v * 8, cpg = C / p.groups;
  // y = a*x + b with a = rstd*gamma, b = beta - mean*a (the form torch's own GroupNorm kernel evaluates)
  float a[8], b[8];
  {
    const uint4 gu = __ldg(reinterpret_cast<const uint4*>(p.gamma + c0));
    const uint4 bu = __ldg(reinterpret_cast<const uint4*>(p.beta + c0));
    const uint32_t ga[4] = {gu.x, gu.y, gu.z, gu.w}, ba[4] = {bu.x, bu.y, bu.z, bu.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 gf = unpack_bf16x2(ga[k]), bf = unpack_bf16x2(ba[k]);
      const float gs[2] = {gf.x, gf.y}, bs[2] = {bf.x, bf.y};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int g = (c0 + 2 * k + e) / cpg;
        const float mean = __ldg(p.mean_rstd + g), rstd = __ldg(p.mean_rstd + p.groups + g);
        a[2 * k + e] = rstd * gs[e];
        b[2 * k + e] = fmaf(-mean, a[2 * k + e], bs[e]);
      }
    }
  }
  const int w_chunk = (p.W + gridDim.y - 1) / gridDim.y;
  const int w_begin = blockIdx.y * w_chunk, w_end = min(p.W, w_begin + w_chunk);
  const int rows = p.T * p.H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int t = row / p.H, h = row - t * p.H;
    const long long zrow = (static_cast<long long>(p.tz_of_t[t]) * p.Hz + (h >> p.shift)) * p.Wz;
    const __nv_bfloat16* xrow = p.x + static_cast<long long>(row) * p.W * C + c0;
    __nv_bfloat16* orow = p.out + static_cast<long long>(row) * p.W * C + c0;
    const __nv_bfloat16* yrow = p.y_lat + zrow * p.ld_lat + c0;
    const __nv_bfloat16* brow = p.b_lat + zrow * p.ld_lat + c0;
#pragma unroll 2
    for (int w = w_begin + slot; w < w_end; w += ppi) {
      const long long zoff = static_cast<long long>(w >> p.shift) * p.ld_lat;
      const uint4 xu = *reinterpret_cast<const uint4*>(xrow + static_cast<long long>(w) * C);
      const uint4 yu = __ldg(reinterpret_cast<const uint4*>(yrow + zoff));
      const uint4 zu = __ldg(reinterpret_cast<const uint4*>(brow + zoff));
      const uint32_t xa[4] = {xu.x, xu.y, xu.z, xu.w}, ya[4] = {yu.x, yu.y, yu.z, yu.w}, za[4] = {zu.x, zu.y, zu.z, zu.w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // roundings where eager bf16 has them: GroupNorm output, * conv_y, + conv_b, SiLU (packed bf16x2 mul / add
        // with explicit .rn so they are not contracted into one fma)
        const float2 xf = unpack_bf16x2(xa[k]);
        const uint32_t n1 = pack_bf16x2(fmaf(xf.x, a[2 * k], b[2 * k]), fmaf(xf.y, a[2 * k + 1], b[2 * k + 1]));
        __nv_bfloat162 n = *reinterpret_cast<const __nv_bfloat162*>(&n1);
        n = __hmul2_rn(n, *reinterpret_cast<const __nv_bfloat162*>(&ya[k]));
        n = __hadd2_rn(n, *reinterpret_cast<const __nv_bfloat162*>(&za[k]));
        uint32_t r = *reinterpret_cast<const uint32_t*>(&n);
        if (p.silu) {
          const float2 f = unpack_bf16x2(r);
          r = pack_bf16x2(__fdividef(f.x, 1.0f + __expf(-f.x)), __fdividef(f.y, 1.0f + __expf(-f.y)));
        }
        o[k] = r;
      }
      *reinterpret_cast<uint4*>(orow + static_cast<long long>(w) * C) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// grid for the row-walking kernels: one block per row up to 148 x 8 blocks; rows are split along W when few
inline dim3 row_grid(int rows, int W, int ppi) {
  const int target = 148 * 4;
  int gx = rows < 148 * 8 ? rows : 148 * 8;
  int gy = 1;
  if (rows < target) {
    gy = (target + rows - 1) / rows;
    const int max_split = (W + ppi - 1) / ppi;         // at least one loop trip per block
    if (gy > max_split) gy = max_split;
    if (gy < 1) gy = 1;
  }
  return dim3(static_cast<unsigned>(gx), static_cast<unsigned>(gy), 1);
}

// ---------------------------------------------------------------------------------------------- nearest upsample x2
struct UpParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  int T_out, H, W, C;          // H, W of the INPUT
  int t_src[16];
};

__global__ void __launch_bounds__(256)
upsample_nearest_kernel(UpParams p) {
  const int C = p.C, vpp = C >> 3, ppi = 256 / vpp;
  const int v = threadIdx.x % vpp, slot = threadIdx.x / vpp;
  if (slot >= ppi) return;
  const int Ho = 2 * p.H, Wo = 2 * p.W;
  const int w_chunk = (Wo + gridDim.y - 1) / gridDim.y;
  const int w_begin = blockIdx.y * w_chunk, w_end = min(Wo, w_begin + w_chunk);
  const int rows = p.T_out * Ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int t = row / Ho, h = row - t * Ho;
    const __nv_bfloat16* srow = p.x + (static_cast<long long>(p.t_src[t]) * p.H + (h >> 1)) * p.W * C + v * 8;
    __nv_bfloat16* orow = p.out + static_cast<long long>(row) * Wo * C + v * 8;
#pragma unroll 4
    for (int w = w_begin + slot; w < w_end; w += ppi)
      *reinterpret_cast<uint4*>(orow + static_cast<long long>(w) * C) =
          __ldg(reinterpret_cast<const uint4*>(srow + static_cast<long long>(w >> 1) * C));
  }
}

// ---------------------------------------------------------------------------------------------- tile blend + crop
struct ComposeParams {
  const __nv_bfloat16* tiles[16];
  int rows, cols;
  int th[4], tw[4];
  int T, H, W, ldc;
  int ev, eh, limit_h, limit_w;
  __nv_bfloat16* out;
};

__device__ __forceinline__ float tile_at(const ComposeParams& p, int i, int j, int t, int y, int x, int c) {
  const __nv_bfloat16* base = p.tiles[i * p.cols + j];
  return __bfloat162float(base[((static_cast<long long>(t) * p.th[i] + y) * p.tw[j] + x) * p.ldc + c]);
}
// b <- a * (1 - k/extent) + b * (k/extent) with the bf16 roundings of the in-place torch expression
__device__ __forceinline__ float blend(float a, float b, int k, int extent) {
  const float wb = static_cast<float>(static_cast<double>(k) / extent);        // python float -> fp32 opmath scalar
  const float wa = static_cast<float>(1.0 - static_cast<double>(k) / extent);
  return bf16_round(bf16_round(a * wa) + bf16_round(b * wb));
}

__global__ void __launch_bounds__(256)
compose_tiles_kernel(ComposeParams p) {
  const long long n = 3ll * p.T * p.H * p.W;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % p.W);
    long long r = idx / p.W;
    const int Y = static_cast<int>(r % p.H);
    r /= p.H;
    const int t = static_cast<int>(r % p.T);
    const int c = static_cast<int>(r / p.T);
    int i = Y / p.limit_h; if (i > p.rows - 1) i = p.rows - 1;
    int j = X / p.limit_w; if (j > p.cols - 1) j = p.cols - 1;
    const int y = Y - i * p.limit_h, x = X - j * p.limit_w;
    // extents as torch's blend_v / blend_h clamp them
    const int ev = i > 0 ? min(min(p.th[i - 1], p.th[i]), p.ev) : 0;
    const int eh = j > 0 ? min(min(p.tw[j - 1], p.tw[j]), p.eh) : 0;
    float val = tile_at(p, i, j, t, y, x, c);
    if (i > 0 && y < ev) {
      // tile above, bottom rows, already blended with ITS left neighbour where x < eh
      const int ya = p.th[i - 1] - ev + y;
      float up = tile_at(p, i - 1, j, t, ya, x, c);
      if (j > 0 && x < eh) up = blend(tile_at(p, i - 1, j - 1, t, ya, p.tw[j - 1] - eh + x, c), up, x, eh);
      val = blend(up, val, y, ev);
    }
    if (j > 0 && x < eh) {
      // left tile, right columns, already blended with ITS upper neighbour where y < ev
      const int xl = p.tw[j - 1] - eh + x;
      float left = tile_at(p, i, j - 1, t, y, xl, c);
      if (i > 0 && y < ev) left = blend(tile_at(p, i - 1, j - 1, t, p.th[i - 1] - ev + y, xl, c), left, y, ev);
      val = blend(left, val, x, eh);
    }
    p.out[idx] = __float2bfloat16_rn(val);
  }
}

inline int grid_for(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace
}  // namespace vgpa

extern "C" size_t vgpa_groupnorm_workspace_bytes(int C) {
  return static_cast<size_t>(vgpa::GN_MAX_BLOCKS) * 2 * static_cast<size_t>(C > 0 ? C : 0) * sizeof(float);
}

namespace vgpa {
int launch_gn_finalize(const float* partial, int nblocks, int C, int groups, double count, float eps, float* mean_rstd, cudaStream_t stream) {
  gn_finalize_kernel<<<groups, 256, 0, stream>>>(partial, nblocks, C, groups, count, eps, mean_rstd);
  VGPA_LAUNCH_CHECK("gn_finalize_kernel");
  return 0;
}
}  // namespace vgpa

extern "C" int vgpa_groupnorm_stats_bf16(const void* x, int64_t n_pixels, int C, int groups, float eps, void* workspace,
                                         size_t workspace_bytes, float* mean_rstd, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(x && workspace && mean_rstd, "vgpa_groupnorm_stats_bf16: null pointer");
  VGPA_CHECK(n_pixels > 0 && C > 0 && C % 8 == 0 && C <= 2048, "vgpa_groupnorm_stats_bf16: bad shape n=%lld C=%d", (long long)n_pixels, C);
  VGPA_CHECK(groups > 0 && groups <= 64 && C % groups == 0, "vgpa_groupnorm_stats_bf16: C=%d not divisible into %d groups", C, groups);
  VGPA_CHECK(GN_THREADS % (C / 8) == 0 || C / 8 > GN_THREADS, "vgpa_groupnorm_stats_bf16: C/8 must divide %d", GN_THREADS);
  VGPA_CHECK(C / 8 <= GN_THREADS, "vgpa_groupnorm_stats_bf16: C=%d too large", C);
  VGPA_CHECK(workspace_bytes >= vgpa_groupnorm_workspace_bytes(C), "vgpa_groupnorm_stats_bf16: workspace too small");
  const int pix_per_iter = GN_THREADS / (C / 8);
  long long nb = (n_pixels + pix_per_iter - 1) / pix_per_iter;
  if (nb > GN_MAX_BLOCKS) nb = GN_MAX_BLOCKS;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  gn_partial_kernel<<<static_cast<int>(nb), GN_THREADS, 2 * C * sizeof(float), s>>>(
      static_cast<const __nv_bfloat16*>(x), n_pixels, C, static_cast<float*>(workspace));
  VGPA_LAUNCH_CHECK("gn_partial_kernel");
  gn_finalize_kernel<<<groups, 256, 0, s>>>(static_cast<const float*>(workspace), static_cast<int>(nb), C, groups,
                                      static_cast<double>(n_pixels) * (C / groups), eps, mean_rstd);
  VGPA_LAUNCH_CHECK("gn_finalize_kernel");
  return 0;
}

extern "C" int vgpa_spatialnorm_apply_bf16(const vgpa_spatialnorm_args* a, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr, "vgpa_spatialnorm_apply_bf16: null args");
  VGPA_CHECK(a->x && a->out && a->mean_rstd && a->gamma && a->beta && a->y_lat && a->b_lat, "vgpa_spatialnorm_apply_bf16: null pointer");
  VGPA_CHECK(a->T > 0 && a->T <= 16 && a->H > 0 && a->W > 0 && a->C > 0 && a->C % 8 == 0 && a->C <= 2048, "vgpa_spatialnorm_apply_bf16: bad shape");
  VGPA_CHECK(a->groups > 0 && a->C % a->groups == 0, "vgpa_spatialnorm_apply_bf16: bad groups");
  VGPA_CHECK(a->ld_lat >= a->C && a->ld_lat % 8 == 0, "vgpa_spatialnorm_apply_bf16: ld_lat=%lld invalid", (long long)a->ld_lat);
  VGPA_CHECK(a->shift >= 0 && ((a->H - 1) >> a->shift) < a->Hz && ((a->W - 1) >> a->shift) < a->Wz,
             "vgpa_spatialnorm_apply_bf16: latent grid %dx%d too small for %dx%d >> %d", a->Hz, a->Wz, a->H, a->W, a->shift);
  SnParams p;
  p.x = static_cast<const __nv_bfloat16*>(a->x);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.mean_rstd = a->mean_rstd;
  p.gamma = static_cast<const __nv_bfloat16*>(a->gamma);
  p.beta = static_cast<const __nv_bfloat16*>(a->beta);
  p.y_lat = static_cast<const __nv_bfloat16*>(a->y_lat);
  p.b_lat = static_cast<const __nv_bfloat16*>(a->b_lat);
  p.ld_lat = a->ld_lat;
  p.T = a->T; p.H = a->H; p.W = a->W; p.C = a->C; p.groups = a->groups;
  p.Hz = a->Hz; p.Wz = a->Wz; p.shift = a->shift; p.silu = a->silu;
  for (int i = 0; i < 16; ++i) p.tz_of_t[i] = a->tz_of_t[i];
  spatialnorm_apply_kernel<<<row_grid(a->T * a->H, a->W, 256 / (a->C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  VGPA_LAUNCH_CHECK("spatialnorm_apply_kernel");
  return 0;
}

extern "C" int vgpa_upsample_nearest_bf16(const void* x, void* out, int T_out, int H, int W, int C, const int32_t* t_src_host,
                                          void* stream) {
  using namespace vgpa;
  VGPA_CHECK(x && out && t_src_host, "vgpa_upsample_nearest_bf16: null pointer");
  VGPA_CHECK(T_out > 0 && T_out <= 16 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && C <= 2048, "vgpa_upsample_nearest_bf16: bad shape");
  UpParams p;
  p.x = static_cast<const __nv_bfloat16*>(x);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.T_out = T_out; p.H = H; p.W = W; p.C = C;
  for (int i = 0; i < 16; ++i) p.t_src[i] = i < T_out ? t_src_host[i] : 0;
  upsample_nearest_kernel<<<row_grid(T_out * 2 * H, 2 * W, 256 / (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  VGPA_LAUNCH_CHECK("upsample_nearest_kernel");
  return 0;
}

extern "C" int vgpa_vae_compose_tiles_bf16(const vgpa_compose_args* a, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr && a->out != nullptr, "vgpa_vae_compose_tiles_bf16: null args");
  VGPA_CHECK(a->rows >= 1 && a->rows <= 4 && a->cols >= 1 && a->cols <= 4, "vgpa_vae_compose_tiles_bf16: at most 4x4 tiles");
  VGPA_CHECK(a->T > 0 && a->H > 0 && a->W > 0 && a->ldc >= 3, "vgpa_vae_compose_tiles_bf16: bad shape");
  ComposeParams p;
  int hsum = 0, wsum = 0;
  for (int i = 0; i < a->rows; ++i) {
    p.th[i] = a->th[i];
    VGPA_CHECK(i == 0 || a->th[i] >= 1, "vgpa_vae_compose_tiles_bf16: empty tile row");
    hsum += (i == a->rows - 1) ? a->th[i] : (a->th[i] < a->limit_h ? a->th[i] : a->limit_h);
  }
  for (int j = 0; j < a->cols; ++j) {
    p.tw[j] = a->tw[j];
    wsum += (j == a->cols - 1) ? a->tw[j] : (a->tw[j] < a->limit_w ? a->tw[j] : a->limit_w);
  }
  VGPA_CHECK(hsum >= a->H && wsum >= a->W, "vgpa_vae_compose_tiles_bf16: tiles (%d x %d after crop) do not cover the %d x %d output",
             hsum, wsum, a->H, a->W);
  for (int i = 0; i + 1 < a->rows; ++i)
    VGPA_CHECK(a->th[i] - a->blend_h >= a->blend_h || a->blend_h == 0, "vgpa_vae_compose_tiles_bf16: tile height %d < 2 x blend extent %d", a->th[i], a->blend_h);
  for (int j = 0; j + 1 < a->cols; ++j)
    VGPA_CHECK(a->tw[j] - a->blend_w >= a->blend_w || a->blend_w == 0, "vgpa_vae_compose_tiles_bf16: tile width %d < 2 x blend extent %d", a->tw[j], a->blend_w);
  for (int k = 0; k < a->rows * a->cols; ++k) {
    VGPA_CHECK(a->tiles[k] != nullptr, "vgpa_vae_compose_tiles_bf16: tile %d is null", k);
    p.tiles[k] = static_cast<const __nv_bfloat16*>(a->tiles[k]);
  }
  p.rows = a->rows; p.cols = a->cols;
  p.T = a->T; p.H = a->H; p.W = a->W; p.ldc = a->ldc;
  p.ev = a->blend_h; p.eh = a->blend_w; p.limit_h = a->limit_h; p.limit_w = a->limit_w;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  const long long n = 3ll * a->T * a->H * a->W;
  compose_tiles_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  VGPA_LAUNCH_CHECK("compose_tiles_kernel");
  return 0;
}
