// K4a: causal 3-D convolution (3x3x3 / 1x3x3) of the CogVideoX VAE decoder as an implicit GEMM on tcgen05.
//
// Replaces the cuDNN conv3d / conv2d calls behind diffusers' CogVideoXCausalConv3d and the Conv2d of
// CogVideoXUpsample3D (reached from AutoencoderKLCogVideoX.decode,
// generate/CogVideoX-5B.py:20-21,72-77).
//
//   out[t,h,w,:] = bias + residual[t,h,w,:] + sum over taps (kt,kh,kw) of x[t+kt, h+kh-1, w+kw-1, :] . W_tap^T
//
// Activations are channels-last [T(+KT-1), H, W, C] bf16, so one output tile of 128 pixels (8 rows x 16
// columns of one frame) against 64 input channels of one tap is a 4-D TMA box {64, 16, 8, 1} whose smem image
// is exactly a 128x64 K-major, 128B-swizzled UMMA A tile. The tap offset is just a coordinate shift; the TMA
// unit zero-fills out-of-range rows/columns, which IS the conv's spatial zero padding. The causal time padding
// (conv_cache or replicated first frame) is materialised by the caller as KT-1 leading frames.
// GEMM K runs over (tap, ci) = KT*9*Cin; weights are stored [Cout_pad, K] in the same order.
//
// Kernel shape = the DiT GEMM (gemm_sm100.cu): persistent CTAs, warp 0 TMA producer, warp 1 tcgen05 issuer,
// warps 2-5 epilogue, multi-stage smem ring, two TMEM accumulators so the epilogue of tile i overlaps tile i+1.
#include "sm100.cuh"
#include "../../include/videogpa_b200.h"

namespace vgpa {
namespace {

constexpr int CV_BM = 128;      // pixels per tile
constexpr int CV_TW = 16;       // tile width  (pixels)
constexpr int CV_TH = 8;        // tile height (pixels)
constexpr int CV_BK = 64;
constexpr int CV_THREADS = 192;
constexpr int CV_MAX_STATS_C = 512;   // widest layer whose GroupNorm statistics the epilogue can accumulate

template <int BN>
struct ConvCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 16 ? 8 : 6);
  static constexpr uint32_t A_BYTES = CV_BM * CV_BK * 2;
  static constexpr uint32_t B_BYTES = BN * CV_BK * 2;
  static constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr uint32_t STATS_BYTES = 4 * 2 * CV_MAX_STATS_C * 4;   // per epilogue warp: [2][C] channel sums / sums of squares
  static constexpr uint32_t SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 + 256 + STATS_BYTES;
};

struct ConvParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;
  int T, H, W, Cin, Cout, KT;
  int ldo, ld_res;
  int tiles_h, tiles_w, num_n;
  float* stats_partial;   // [gridDim.x][2][Cout] per-CTA channel sums / sums of squares of the STORED (bf16) output, or NULL
};

// Sum over the 32 lanes of a warp of 16 per-lane values, 16 shuffles instead of 80: every step exchanges the half a lane does
// not keep. Afterwards lane l holds the total of value index ((l >> 1) & 15) (lanes l and l ^ 1 hold the same total).
__device__ __forceinline__ float warp_transpose_sum16(const float (&a)[16], int lane) {
  float b[8], c[4], d[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = h16 ? a[8 + i] : a[i], send = h16 ? a[i] : a[8 + i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = h8 ? b[4 + i] : b[i], send = h8 ? b[i] : b[4 + i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = h4 ? c[2 + i] : c[i], send = h4 ? c[i] : c[2 + i];
    d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = h2 ? d[1] : d[0], send = h2 ? d[0] : d[1];
  float e = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  e += __shfl_xor_sync(0xffffffffu, e, 1);
  return e;
}

template <int BN>
__global__ void __launch_bounds__(CV_THREADS, 1)
vae_conv3d_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvParams p) {
  using Cfg = ConvCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_stats = reinterpret_cast<float*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES) + 256);   // [4][2][Cout]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.stats_partial != nullptr)
    for (int i = threadIdx.x; i < 4 * 2 * p.Cout; i += CV_THREADS) s_stats[i] = 0.f;
  const int tiles_per_frame = p.tiles_h * p.tiles_w;
  const int num_m = p.T * tiles_per_frame;
  const int num_tiles = num_m * p.num_n;
  const int cblocks = p.Cin / CV_BK;
  const int nk = p.KT * 9 * cblocks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (frame t, first row h0, first column w0, n block); the n blocks of one pixel tile are adjacent so
  // CTAs running side by side share the activation tile through L2
  auto tile_coords = [&](int tile, int& t, int& h0, int& w0, int& n_blk) {
    n_blk = tile % p.num_n;
    const int m = tile / p.num_n;
    t = m / tiles_per_frame;
    const int r = m - t * tiles_per_frame;
    h0 = (r / p.tiles_w) * CV_TH;
    w0 = (r % p.tiles_w) * CV_TW;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int t, h0, w0, n_blk;
        tile_coords(tile, t, h0, w0, n_blk);
        int kb = 0;
        for (int tap = 0; tap < p.KT * 9; ++tap) {
          const int kt = tap / 9, kh = (tap % 9) / 3, kw = tap % 3;
          for (int cb = 0; cb < cblocks; ++cb, ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
            ptx::tma_load_4d(sA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], cb * CV_BK, w0 + kw - 1, h0 + kh - 1, t + kt);
            ptx::tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * CV_BK, n_blk * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::idesc_bf16(CV_BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < nk; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t adesc = ptx::smem_desc_sw128(ptx::smem_u32(sA + stage * Cfg::A_BYTES), 16, 1024);
          const uint64_t bdesc = ptx::smem_desc_sw128(ptx::smem_u32(sB + stage * Cfg::B_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < CV_BK / 16; ++k)
            ptx::umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&empty_bar[stage]);
          if (kb == nk - 1) ptx::umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps: TMEM -> bias/residual -> bf16
    const int quarter = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int t, h0, w0, n_blk;
      tile_coords(tile, t, h0, w0, n_blk);
      ptx::mbar_wait(&tfull_bar[as], aphase);
      ptx::tc_fence_after();
      const int r = quarter * 32 + lane;
      const int h = h0 + (r >> 4), w = w0 + (r & 15);
      const bool live = h < p.H && w < p.W;
      const long long pix = (static_cast<long long>(t) * p.H + (live ? h : 0)) * p.W + (live ? w : 0);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t acc[16];
        ptx::tmem_ld_32x16(t_row + c0, acc);
        ptx::tmem_ld_wait();
        const int col = n_blk * BN + c0;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
        if (live && col < p.Cout) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[i]);
          if (p.bias != nullptr) {
            const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const uint4 b = __ldg(bp + q);
              const uint32_t bu[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) { const float2 f = unpack_bf16x2(bu[i]); v[q * 8 + 2 * i] += f.x; v[q * 8 + 2 * i + 1] += f.y; }
            }
          }
          if (p.residual != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.ld_res + col);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const uint4 b = rp[q];
              const uint32_t bu[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_bf16x2(bu[i]);
                // eager bf16: conv output is rounded to bf16 before the residual add
                v[q * 8 + 2 * i] = f.x + bf16_round(v[q * 8 + 2 * i]);
                v[q * 8 + 2 * i + 1] = f.y + bf16_round(v[q * 8 + 2 * i + 1]);
              }
            }
          }
          uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.ldo + col);
          uint4 o0, o1;
          o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);   o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
          o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]); o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
          op[0] = o0;
          op[1] = o1;
        }
        if (p.stats_partial != nullptr && col < p.Cout) {        // warp-uniform: every lane takes part in the shuffles
          // GroupNorm statistics of the tensor being written (what the next SpatialNorm normalises with): per-channel sum and
          // sum of squares of the bf16 values as stored; pixels outside the image contribute zeros. Each epilogue warp owns
          // a [2][Cout] slice in shared memory, added to in tile order: deterministic.
          float sq[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] = bf16_round(v[i]); sq[i] = v[i] * v[i]; }
          const float ts = warp_transpose_sum16(v, lane), tq = warp_transpose_sum16(sq, lane);
          if ((lane & 1) == 0) {
            float* st = s_stats + quarter * 2 * p.Cout + col + (lane >> 1);
            st[0] += ts;
            st[p.Cout] += tq;
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (p.stats_partial != nullptr) {
    const int n2 = 2 * p.Cout;
    for (int i = threadIdx.x; i < n2; i += CV_THREADS)
      p.stats_partial[static_cast<long long>(blockIdx.x) * n2 + i] = ((s_stats[i] + s_stats[n2 + i]) + s_stats[2 * n2 + i]) + s_stats[3 * n2 + i];
  }
}

struct GnOut { float* mean_rstd; int groups; float eps; };

template <int BN>
int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, const GnOut& gn, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    VGPA_CUDA(cudaFuncSetAttribute(vae_conv3d_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const long long num_tiles = static_cast<long long>(p.T) * p.tiles_h * p.tiles_w * p.num_n;
  const int grid = num_tiles < num_sms() ? static_cast<int>(num_tiles) : num_sms();
  vae_conv3d_kernel<BN><<<grid, CV_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  VGPA_LAUNCH_CHECK("vae_conv3d_kernel");
  if (p.stats_partial != nullptr) {
    // fold the per-CTA partials in fp64 in a fixed order and reduce channels to groups (vae_norm.cu)
    return launch_gn_finalize(p.stats_partial, grid, p.Cout, gn.groups, static_cast<double>(p.T) * p.H * p.W * (p.Cout / gn.groups), gn.eps,
                              gn.mean_rstd, stream);
  }
  return 0;
}

}  // namespace
}  // namespace vgpa

extern "C" size_t vgpa_conv3d_gn_workspace_bytes(int Cout) {
  return static_cast<size_t>(vgpa::num_sms() > 0 ? vgpa::num_sms() : 148) * 2 * (Cout > 0 ? Cout : 0) * sizeof(float) + 256;
}

extern "C" int vgpa_conv3d_causal_bf16(const vgpa_conv3d_args* a, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr, "vgpa_conv3d_causal_bf16: null args");
  VGPA_CHECK(a->x && a->w && a->out, "vgpa_conv3d_causal_bf16: null tensor pointer");
  VGPA_CHECK(a->T > 0 && a->H > 0 && a->W > 0, "vgpa_conv3d_causal_bf16: bad shape T=%d H=%d W=%d", a->T, a->H, a->W);
  VGPA_CHECK(a->KT == 1 || a->KT == 3, "vgpa_conv3d_causal_bf16: KT must be 1 or 3 (got %d)", a->KT);
  VGPA_CHECK(a->Cin > 0 && a->Cin % 64 == 0, "vgpa_conv3d_causal_bf16: Cin=%d must be a multiple of 64", a->Cin);
  VGPA_CHECK(a->Cout > 0 && a->Cout <= a->Cout_pad, "vgpa_conv3d_causal_bf16: Cout=%d Cout_pad=%d", a->Cout, a->Cout_pad);
  int BN = 0;
  if (a->Cout_pad % 256 == 0) BN = 256;
  else if (a->Cout_pad == 128) BN = 128;
  else if (a->Cout_pad == 64) BN = 64;
  else if (a->Cout_pad == 16) BN = 16;
  VGPA_CHECK(BN != 0, "vgpa_conv3d_causal_bf16: Cout_pad=%d must be 16, 64, 128 or a multiple of 256", a->Cout_pad);
  if (BN == 256) {
    // low-resolution layers have fewer 128x256 output tiles than SMs (30x45x2 frames, 512 channels: 48): halve the
    // N tile when that shortens the schedule. A 128x128 tile costs ~0.58 of a 128x256 one (its MMAs are bound by the
    // shared-memory operand reads, not the tensor pipe).
    const long long m_tiles = static_cast<long long>(a->T) * ((a->H + CV_TH - 1) / CV_TH) * ((a->W + CV_TW - 1) / CV_TW);
    const long long it256 = m_tiles * (a->Cout_pad / 256), it128 = 2 * it256;
    const double c256 = static_cast<double>((it256 + 147) / 148), c128 = 0.58 * static_cast<double>((it128 + 147) / 148);
    if (c128 < c256) BN = 128;
  }
  VGPA_CHECK(a->Cout % 16 == 0 || a->Cout_pad == 16, "vgpa_conv3d_causal_bf16: Cout=%d must be a multiple of 16 (or Cout_pad 16)", a->Cout);
  VGPA_CHECK(a->ldo % 8 == 0 && a->ldo >= (a->Cout_pad == 16 ? 16 : a->Cout), "vgpa_conv3d_causal_bf16: ldo=%d invalid", a->ldo);
  VGPA_CHECK(a->residual == nullptr || (a->ld_res % 8 == 0 && a->ld_res >= a->Cout), "vgpa_conv3d_causal_bf16: ld_res=%d invalid", a->ld_res);
  VGPA_CHECK(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w) | reinterpret_cast<uintptr_t>(a->out) |
               reinterpret_cast<uintptr_t>(a->residual) | reinterpret_cast<uintptr_t>(a->bias)) & 15) == 0,
             "vgpa_conv3d_causal_bf16: pointers must be 16-byte aligned");

  const uint64_t K = static_cast<uint64_t>(a->KT) * 9 * a->Cin;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)(a->T + a->KT - 1)};
    const uint64_t str[3] = {(uint64_t)a->Cin * 2, (uint64_t)a->W * a->Cin * 2, (uint64_t)a->H * a->W * a->Cin * 2};
    const uint32_t box[4] = {CV_BK, CV_TW, CV_TH, 1};
    if (int rc = make_tmap_bf16(&tmA, a->x, 4, dims, str, box)) return rc;
  }
  {
    const uint64_t dims[2] = {K, (uint64_t)a->Cout_pad};
    const uint64_t str[1] = {K * 2};
    const uint32_t box[2] = {CV_BK, (uint32_t)BN};
    if (int rc = make_tmap_bf16(&tmB, a->w, 2, dims, str, box)) return rc;
  }
  ConvParams p;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.bias = static_cast<const __nv_bfloat16*>(a->bias);
  p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  p.T = a->T; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.KT = a->KT;
  p.Cout = (a->Cout_pad == 16) ? 16 : a->Cout;   // the 16-wide (conv_out) variant stores all 16 padded channels
  p.ldo = a->ldo; p.ld_res = a->ld_res;
  p.tiles_h = (a->H + CV_TH - 1) / CV_TH;
  p.tiles_w = (a->W + CV_TW - 1) / CV_TW;
  p.num_n = a->Cout_pad / BN;
  p.stats_partial = nullptr;
  GnOut gn{a->gn_mean_rstd, a->gn_groups, a->gn_eps};
  if (a->gn_mean_rstd != nullptr) {
    VGPA_CHECK(a->gn_workspace != nullptr && (reinterpret_cast<uintptr_t>(a->gn_workspace) & 15) == 0, "vgpa_conv3d_causal_bf16: gn_workspace missing or misaligned");
    VGPA_CHECK(a->Cout_pad != 16 && a->Cout <= CV_MAX_STATS_C && a->gn_groups > 0 && a->gn_groups <= 64 && a->Cout % a->gn_groups == 0,
               "vgpa_conv3d_causal_bf16: fused GroupNorm statistics need Cout <= %d divisible into gn_groups (Cout=%d groups=%d)", CV_MAX_STATS_C,
               a->Cout, a->gn_groups);
    p.stats_partial = static_cast<float*>(a->gn_workspace);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (BN) {
    case 256: return launch_conv<256>(tmA, tmB, p, gn, s);
    case 128: return launch_conv<128>(tmA, tmB, p, gn, s);
    case 64: return launch_conv<64>(tmA, tmB, p, gn, s);
    default: return launch_conv<16>(tmA, tmB, p, gn, s);
  }
}
