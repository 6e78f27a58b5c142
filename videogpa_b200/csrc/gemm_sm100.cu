// K2: persistent TMA + tcgen05 GEMM for the DiT linears, out = epilogue(A[M,K] . W[N,K]^T).
//
// Replaces the cuBLAS calls behind diffusers' to_q/to_k/to_v/to_out Linear and FeedForward
// (reference call site generate/CogVideoX-5B.py:72).
//
// Shape of the kernel (one CTA per SM, 192 threads):
//   warp 0      TMA producer: A tile 128x64 and W tile BNx64 (bf16, 128B swizzle) into a smem ring
//   warp 1      MMA issuer: lane 0 issues tcgen05.mma 128xBNx16 into one of two TMEM accumulators
//   warps 2..5  epilogue: tcgen05.ld -> registers -> fused epilogue -> bf16 global stores
// The two TMEM accumulator stages let the epilogue of tile i overlap the mainloop of tile i+1.
#include "sm100.cuh"
#include "../../include/videogpa_b200.h"
#include <stdlib.h>

namespace vgpa {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int GROUP_M_DEFAULT = 16;

struct EpiParams {
  __nv_bfloat16* out;
  int ldo;
  const __nv_bfloat16* bias;
  // gated residual
  const __nv_bfloat16* gate_txt;
  const __nv_bfloat16* gate_vid;
  long long gate_stride_b;
  int rows_per_sample;
  int text_rows;
  // qkv
  const float* ln_q_w;
  const float* ln_q_b;
  const float* ln_k_w;
  const float* ln_k_b;
  float ln_eps;
  const float* rope_cos;
  const float* rope_sin;
  int model_dim;
  float alpha;
  // rasterisation / L2 policy (host-chosen per launch)
  int group_m;             // M tiles (tile pairs with a cluster) per raster group
  int stream_out;          // 1: output stores carry the evict-first (streaming) hint
  uint64_t hint_a, hint_w; // L2 eviction policy of the A / W operand loads
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN > 128 ? 5 : 6);
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);   // allocations are powers of two
  static constexpr uint32_t SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align*/ + 256;
};

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))), tanh(u) = 1 - 2 / (exp(2u) + 1)
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float e = __expf(2.0f * u);
  const float t = 1.0f - __fdividef(2.0f, e + 1.0f);
  return 0.5f * x * (1.0f + t);
}

// Epilogue over one 64-column group held in v[64] for output row `row` (may be >= M: no stores). NC = 32: only the first 32
// columns are valid (the tail group of a 160-wide tile).
template <int EPI, int NC = 64>
__device__ __forceinline__ void epilogue_group(float (&v)[64], int row, int col0, int M,
                                               const EpiParams& ep, const uint4 (&res)[8]) {
  static_assert(NC == 64 || (EPI != VGPA_EPI_QKV && EPI != VGPA_EPI_GATE_RES_F32), "the per-head QKV / fp32-residual epilogues work on whole 64-column groups");
  const bool live = row < M;
  if (ep.bias != nullptr) {
    const uint4* bp = reinterpret_cast<const uint4*>(ep.bias + col0);
#pragma unroll
    for (int i = 0; i < NC / 8; ++i) {
      const uint4 b = __ldg(bp + i);
      const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z),
                   b3 = unpack_bf16x2(b.w);
      v[i * 8 + 0] += b0.x; v[i * 8 + 1] += b0.y; v[i * 8 + 2] += b1.x; v[i * 8 + 3] += b1.y;
      v[i * 8 + 4] += b2.x; v[i * 8 + 5] += b2.y; v[i * 8 + 6] += b3.x; v[i * 8 + 7] += b3.y;
    }
  }
  __nv_bfloat16* orow = ep.out + static_cast<size_t>(live ? row : 0) * ep.ldo + col0;

  if constexpr (EPI == VGPA_EPI_BIAS_GELU) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = gelu_tanh(bf16_round(v[i]));
  } else if constexpr (EPI == VGPA_EPI_GATE_RES) {
    // x <- x + gate * y with eager-bf16 roundings (y, gate*y, sum each rounded to bf16)
    int srow = row, b = 0;
    if (ep.rows_per_sample > 0) { b = row / ep.rows_per_sample; srow = row - b * ep.rows_per_sample; }
    const __nv_bfloat16* g = (srow < ep.text_rows ? ep.gate_txt : ep.gate_vid);
    if (live) {
      const uint4* gp = (g != nullptr) ? reinterpret_cast<const uint4*>(g + b * ep.gate_stride_b + col0) : nullptr;
#pragma unroll
      for (int i = 0; i < NC / 8; ++i) {
        const uint4 r = res[i];                          // residual row chunk, prefetched one column group ahead
        uint4 gg = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);  // bf16 1.0
        if (gp != nullptr) gg = __ldg(gp + i);
        const uint32_t ru[4] = {r.x, r.y, r.z, r.w};
        const uint32_t gu[4] = {gg.x, gg.y, gg.z, gg.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 rr = unpack_bf16x2(ru[j]);
          const float2 gf = unpack_bf16x2(gu[j]);
          const float y0 = bf16_round(v[i * 8 + 2 * j]), y1 = bf16_round(v[i * 8 + 2 * j + 1]);
          v[i * 8 + 2 * j] = rr.x + bf16_round(gf.x * y0);
          v[i * 8 + 2 * j + 1] = rr.y + bf16_round(gf.y * y1);
        }
      }
    }
  } else if constexpr (EPI == VGPA_EPI_GATE_RES_F32) {
    // fp32 residual stream: x <- x + gate * bf16(y), evaluated and stored in fp32 (no bf16 store below)
    int srow = row, b = 0;
    if (ep.rows_per_sample > 0) { b = row / ep.rows_per_sample; srow = row - b * ep.rows_per_sample; }
    const __nv_bfloat16* g = (srow < ep.text_rows ? ep.gate_txt : ep.gate_vid);
    if (live) {
      float4* xo = reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + static_cast<size_t>(row) * ep.ldo + col0);
      const uint2* gp = (g != nullptr) ? reinterpret_cast<const uint2*>(g + b * ep.gate_stride_b + col0) : nullptr;
      float4 r[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = xo[i];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float2 g0 = make_float2(1.f, 1.f), g1 = g0;
        if (gp != nullptr) { const uint2 gg = __ldg(gp + i); g0 = unpack_bf16x2(gg.x); g1 = unpack_bf16x2(gg.y); }
        r[i].x += g0.x * bf16_round(v[4 * i + 0]);
        r[i].y += g0.y * bf16_round(v[4 * i + 1]);
        r[i].z += g1.x * bf16_round(v[4 * i + 2]);
        r[i].w += g1.y * bf16_round(v[4 * i + 3]);
        xo[i] = r[i];
      }
    }
    return;
  } else if constexpr (EPI == VGPA_EPI_ACCUM) {
    // out <- bf16(out + alpha * acc): one rounding (PEFT merge `weight += scaling * B @ A`)
    if (live) {
      const uint4* rp = reinterpret_cast<const uint4*>(orow);
#pragma unroll
      for (int i = 0; i < NC / 8; ++i) {
        const uint4 r = rp[i];
        const uint32_t ru[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 rr = unpack_bf16x2(ru[j]);
          v[i * 8 + 2 * j] = rr.x + ep.alpha * v[i * 8 + 2 * j];
          v[i * 8 + 2 * j + 1] = rr.y + ep.alpha * v[i * 8 + 2 * j + 1];
        }
      }
    }
  } else if constexpr (EPI == VGPA_EPI_QKV) {
    const int which = col0 / ep.model_dim;  // 0 = q, 1 = k, 2 = v
    if (which < 2) {
      // per-head LayerNorm(64) on the bf16-rounded projection, then (video rows) interleaved RoPE
      float mean = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) { v[i] = bf16_round(v[i]); mean += v[i]; }
      mean *= (1.0f / 64.0f);
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) { const float d = v[i] - mean; var += d * d; }
      var *= (1.0f / 64.0f);
      const float rstd = rsqrtf(var + ep.ln_eps);
      const float* w = which == 0 ? ep.ln_q_w : ep.ln_k_w;
      const float* bb = which == 0 ? ep.ln_q_b : ep.ln_k_b;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + i));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bb + i));
        v[i + 0] = bf16_round((v[i + 0] - mean) * rstd * w4.x + b4.x);
        v[i + 1] = bf16_round((v[i + 1] - mean) * rstd * w4.y + b4.y);
        v[i + 2] = bf16_round((v[i + 2] - mean) * rstd * w4.z + b4.z);
        v[i + 3] = bf16_round((v[i + 3] - mean) * rstd * w4.w + b4.w);
      }
      int srow = row;
      if (ep.rows_per_sample > 0) srow = row % ep.rows_per_sample;
      if (ep.rope_cos != nullptr && srow >= ep.text_rows && live) {
        const float* cs = ep.rope_cos + static_cast<size_t>(srow - ep.text_rows) * 64;
        const float* sn = ep.rope_sin + static_cast<size_t>(srow - ep.text_rows) * 64;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs + i));
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(sn + i));
          const float x0 = v[i], x1 = v[i + 1], x2 = v[i + 2], x3 = v[i + 3];
          v[i + 0] = x0 * c4.x + (-x1) * s4.x;
          v[i + 1] = x1 * c4.y + x0 * s4.y;
          v[i + 2] = x2 * c4.z + (-x3) * s4.z;
          v[i + 3] = x3 * c4.w + x2 * s4.w;
        }
      }
    }
  }
  if (live) {
    uint4* op = reinterpret_cast<uint4*>(orow);
#pragma unroll
    for (int i = 0; i < NC / 8; ++i) {
      uint4 o;
      o.x = pack_bf16x2(v[i * 8 + 0], v[i * 8 + 1]);
      o.y = pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
      o.z = pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]);
      o.w = pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
      if (ep.stream_out) __stcs(op + i, o); else op[i] = o;
    }
  }
}

// CL = 2: the two CTAs of a cluster work on vertically adjacent M tiles of the same N block; each loads only half of the
// W tile and multicasts it into both CTAs' shared memory, cutting the L2 -> SM operand traffic from 48 KB to 32 KB per
// k block (ncu showed the MMA warp waiting on TMA data about half of its time with one CTA per tile pair). The stage is
// released by a multicast tcgen05.commit so that a slot is only refilled once BOTH CTAs have consumed it.
template <int BN, int EPI, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 int M, int N, int K, EpiParams ep) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = (CL == 2) ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int num_m = ((M + BM - 1) / BM + CL - 1) / CL;      // M tiles per CTA of the cluster (pairs for CL = 2)
  const int num_n = N / BN;
  const int num_tiles = num_m * num_n;
  const int nk = (K + BK - 1) / BK;
  const int first_tile = blockIdx.x / CL, tile_step = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], CL);                      // one tcgen05.commit per CTA of the cluster
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
  if (CL == 2) ptx::cluster_sync_all();                        // peer barriers are initialised before anything remote lands
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int per_group = ep.group_m * num_n;
    const int group = tile / per_group;
    const int first_m = group * ep.group_m;
    const int gsz = min(num_m - first_m, ep.group_m);
    const int in_group = tile - group * per_group;
    m_blk = (first_m + in_group % gsz) * CL + cta_rank;       // may be one past the last M tile: loads are zero-filled, stores masked
    n_blk = in_group / gsz;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {   // elect.sync keeps TMA / tcgen05 issue free of per-lane waterfall loops
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        for (int kb = 0; kb < nk; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          ptx::tma_load_2d_hint(sA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * BK, m_blk * BM, ep.hint_a);
          if (CL == 2) {
            ptx::tma_load_2d_multicast_hint(sB + stage * Cfg::B_BYTES + cta_rank * (Cfg::B_BYTES / 2), &tmB, &full_bar[stage], kb * BK,
                                            n_blk * BN + cta_rank * (BN / 2), 0x3, ep.hint_w);
          } else {
            ptx::tma_load_2d_hint(sB + stage * Cfg::B_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN, ep.hint_w);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
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
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes per K=16 step inside the 128B swizzle atom (start-address field is >>4)
            ptx::umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (CL == 2) ptx::umma_commit_multicast(&empty_bar[stage], 0x3); else ptx::umma_commit(&empty_bar[stage]);
          if (kb == nk - 1) ptx::umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      const int row = m_blk * BM + quarter * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
      // Gated residual: the residual row chunk of column group g+1 is loaded while group g is processed (and group
      // 0 before the accumulator wait), so the HBM round trip is off the epilogue's critical path. Left to ptxas,
      // each of the eight 16-byte loads was sunk next to its use behind the previous chunk's store, serialising eight
      // round trips per group (+1 ms per GEMM at the DiT shapes).
      uint4 res[8];
      auto load_res = [&](int g, uint4 (&dst)[8]) {
        if constexpr (EPI == VGPA_EPI_GATE_RES) {
          const uint4* rp = reinterpret_cast<const uint4*>(ep.out + static_cast<size_t>(row < M ? row : 0) * ep.ldo + n_blk * BN + g * 64);
          const int nv = (BN - g * 64 >= 64) ? 8 : (BN - g * 64) / 8;      // the tail group of a 160-wide tile has 32 columns
#pragma unroll
          for (int i = 0; i < 8; ++i) if (i < nv) dst[i] = rp[i];
        }
      };
      load_res(0, res);
      ptx::mbar_wait(&tfull_bar[as], aphase);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < BN / 64; ++g) {
        uint4 nxt[8];
        if (g + 1 < (BN + 63) / 64) load_res(g + 1, nxt);
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32(t_row + g * 64, r0);
        ptx::tmem_ld_32x32(t_row + g * 64 + 32, r1);
        ptx::tmem_ld_wait();
        float v[64];
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r0[i]); v[32 + i] = __uint_as_float(r1[i]); }
        epilogue_group<EPI>(v, row, n_blk * BN + g * 64, M, ep, res);
        if constexpr (EPI == VGPA_EPI_GATE_RES) {
#pragma unroll
          for (int i = 0; i < 8; ++i) res[i] = nxt[i];
        }
      }
      if constexpr (BN % 64 != 0) {                       // 32-column tail group of a 160-wide tile
        constexpr int g = BN / 64;
        uint32_t r0[32];
        ptx::tmem_ld_32x32(t_row + g * 64, r0);
        ptx::tmem_ld_wait();
        float v[64];
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r0[i]); v[32 + i] = 0.f; }
        epilogue_group<EPI, 32>(v, row, n_blk * BN + g * 64, M, ep, res);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CL == 2) ptx::cluster_sync_all();                        // no CTA leaves while its peer may still signal its barriers
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, int EPI, int CL>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K,
                const EpiParams& ep, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    VGPA_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, EPI, CL>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int num_units = ((((M + BM - 1) / BM) + CL - 1) / CL) * (N / BN);      // tiles (CL = 1) or tile pairs (CL = 2)
  const int max_units = num_sms() / CL;
  const int grid = (num_units < max_units ? num_units : max_units) * CL;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VGPA_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, EPI, CL>, tmA, tmB, M, N, K, ep));
  return 0;
}

}  // namespace

}  // namespace vgpa

extern "C" int vgpa_linear_bf16(const vgpa_linear_args* a, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr, "vgpa_linear_bf16: null args");
  VGPA_CHECK(a->M > 0 && a->N > 0 && a->K > 0, "vgpa_linear_bf16: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  VGPA_CHECK(a->N % 64 == 0, "vgpa_linear_bf16: N=%d must be a multiple of 64", a->N);
  VGPA_CHECK(a->K % 8 == 0 && a->lda % 8 == 0, "vgpa_linear_bf16: K and lda must be multiples of 8");
  VGPA_CHECK(a->ldo % 8 == 0 && a->ldo >= a->N, "vgpa_linear_bf16: ldo=%d invalid", a->ldo);
  VGPA_CHECK(a->A && a->W && a->out, "vgpa_linear_bf16: null tensor pointer");
  VGPA_CHECK((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->W) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
             "vgpa_linear_bf16: pointers must be 16-byte aligned");
  int BN = (a->N % 256 == 0) ? 256 : 64;
  // Skinny problems (T5 prompt encoder: M = 226) are bound by streaming W once from HBM, and what limits that stream is each SM's
  // ingest from L2 (~106 GB/s measured): a tile moves A (16 KB) + W (BN / 8 KB) per 64-deep k block and only the W part is new
  // data, and a tile count just above the SM count costs a whole second wave (N = 10 240 at BN = 128: 160 tiles on 148 SMs).
  // Pick the tile width that minimises waves x bytes per k block among the widths that divide N; 160 and 192 exist for the
  // epilogues the skinny callers use (bias, bias + GELU, gated residual): 10 240 = 64 x 160 and 12 288 = 64 x 192 are one wave.
  const long long m_tiles = (a->M + BM - 1) / BM;
  if (BN == 256 && m_tiles * (a->N / 256) < 148) {
    static int k_wide = -1;
    if (k_wide < 0) { const char* e = getenv("VGPA_GEMM_SKINNY_TILES"); k_wide = e ? atoi(e) : 1; }   // dev: 0 = 64-wide tiles only
    const bool odd_ok = a->epilogue == VGPA_EPI_BIAS || a->epilogue == VGPA_EPI_BIAS_GELU || a->epilogue == VGPA_EPI_GATE_RES;
    const int cand[5] = {256, 192, 160, 128, 64};
    double best = 1e30;
    BN = 64;
    for (int c = 0; c < 5; ++c) {
      const int bn = cand[c];
      if (a->N % bn != 0 || (!k_wide && bn != 64) || ((bn == 160 || bn == 192) && !odd_ok)) continue;
      const long long tiles = m_tiles * (a->N / bn);
      const long long waves = (tiles + num_sms() - 1) / num_sms();
      const double t = static_cast<double>(waves) * (16.0 + bn / 8.0);        // KB per k block of a tile x tiles an SM runs in sequence
      if (t < best) { best = t; BN = bn; }
    }
  }
  // cluster of 2 with W-tile multicast for the big GEMMs (development knob VGPA_GEMM_CLUSTER=0 turns it off)
  static int use_cluster = -1;
  if (use_cluster < 0) {
    const char* e = getenv("VGPA_GEMM_CLUSTER");
    use_cluster = e ? atoi(e) : 1;
  }
  const int CL = (use_cluster && BN == 256 && (a->M + BM - 1) / BM >= 16) ? 2 : 1;

  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.out = static_cast<__nv_bfloat16*>(a->out);
  ep.ldo = a->ldo;
  ep.bias = static_cast<const __nv_bfloat16*>(a->bias);
  ep.gate_txt = static_cast<const __nv_bfloat16*>(a->gate_txt);
  ep.gate_vid = static_cast<const __nv_bfloat16*>(a->gate_vid);
  ep.gate_stride_b = a->gate_stride_b;
  ep.rows_per_sample = a->rows_per_sample;
  ep.text_rows = a->text_rows;
  ep.ln_q_w = a->ln_q_w; ep.ln_q_b = a->ln_q_b; ep.ln_k_w = a->ln_k_w; ep.ln_k_b = a->ln_k_b;
  ep.ln_eps = a->ln_eps;
  ep.rope_cos = a->rope_cos; ep.rope_sin = a->rope_sin;
  ep.model_dim = a->model_dim;
  ep.alpha = a->alpha;
  // Rasterisation and L2 policy. A raster group is `group_m` M tiles (tile pairs with a cluster) swept across all of N: the A rows of
  // the group are what must survive in L2 while W and the output stream through. Measured on the DiT shapes (profiles/r02_gemm_l2.md):
  // for K = 3072 a 24-pair group (6144 rows, 38 MB) with evict-last A loads and streaming (evict-first) output stores reads 843 MB from
  // DRAM for FF1 against 908 MB with 16 / no hints and is 1.5-2 % faster; 32 pairs (50 MB) no longer survives (1.42 GB). With K = 12288
  // (FF2) a group already exceeds L2 and the hints cost 3 %: plain loads, 16. VGPA_GEMM_GROUP_M / _STREAM_OUT / _HINTS override (dev).
  static int k_group = -2, k_stream = -2, k_hints = -2;
  if (k_group == -2) {
    const char* e = getenv("VGPA_GEMM_GROUP_M");  k_group = e ? atoi(e) : -1;
    e = getenv("VGPA_GEMM_STREAM_OUT");           k_stream = e ? atoi(e) : -1;
    e = getenv("VGPA_GEMM_HINTS");                k_hints = e ? atoi(e) : -1;
  }
  const bool short_k = a->K <= 4096;
  const int hints = k_hints >= 0 ? k_hints : (short_k ? 1 : 0);
  ep.group_m = k_group >= 1 ? k_group : (short_k ? 24 : GROUP_M_DEFAULT);
  // in-place epilogues re-read what they store (residual / accumulate): never stream those
  const bool in_place = a->epilogue == VGPA_EPI_GATE_RES || a->epilogue == VGPA_EPI_ACCUM || a->epilogue == VGPA_EPI_GATE_RES_F32;
  ep.stream_out = ((k_stream >= 0 ? k_stream : (short_k ? 1 : 0)) && !in_place) ? 1 : 0;
  ep.hint_a = (hints & 1) ? ptx::L2_EVICT_LAST : ptx::L2_EVICT_NORMAL;
  ep.hint_w = (hints & 2) ? ptx::L2_EVICT_FIRST : ((hints & 4) ? ptx::L2_EVICT_LAST : ptx::L2_EVICT_NORMAL);
  if (a->epilogue == VGPA_EPI_QKV) {
    VGPA_CHECK(a->model_dim > 0 && a->model_dim % 64 == 0 && a->N == 3 * a->model_dim,
               "vgpa_linear_bf16: QKV epilogue needs N == 3*model_dim (N=%d model_dim=%d)", a->N, a->model_dim);
    VGPA_CHECK(a->ln_q_w && a->ln_q_b && a->ln_k_w && a->ln_k_b, "vgpa_linear_bf16: QKV epilogue needs q/k LayerNorm params");
    VGPA_CHECK((a->rope_cos == nullptr) == (a->rope_sin == nullptr), "vgpa_linear_bf16: rope cos/sin must both be set or both null");
  }
  if (a->epilogue == VGPA_EPI_GATE_RES || a->epilogue == VGPA_EPI_GATE_RES_F32) {
    VGPA_CHECK((a->gate_txt == nullptr) == (a->gate_vid == nullptr), "vgpa_linear_bf16: gate_txt/gate_vid must both be set or both null");
  }

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    const uint64_t strides[1] = {(uint64_t)a->lda * 2};
    const uint32_t box[2] = {BK, BM};
    if (int rc = make_tmap_bf16(&tmA, a->A, 2, dims, strides, box)) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    const uint64_t strides[1] = {(uint64_t)a->K * 2};
    const uint32_t box[2] = {BK, (uint32_t)(BN / CL)};
    if (int rc = make_tmap_bf16(&tmB, a->W, 2, dims, strides, box)) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define VGPA_GEMM_DISPATCH(EPI)                                                          \
  case EPI:                                                                              \
    return BN == 256 ? (CL == 2 ? launch_gemm<256, EPI, 2>(tmA, tmB, a->M, a->N, a->K, ep, s)    \
                                : launch_gemm<256, EPI, 1>(tmA, tmB, a->M, a->N, a->K, ep, s))   \
           : BN == 128 ? launch_gemm<128, EPI, 1>(tmA, tmB, a->M, a->N, a->K, ep, s)             \
                       : launch_gemm<64, EPI, 1>(tmA, tmB, a->M, a->N, a->K, ep, s);
#define VGPA_GEMM_DISPATCH_ODD(EPI)                                                      \
  case EPI:                                                                              \
    return BN == 192 ? launch_gemm<192, EPI, 1>(tmA, tmB, a->M, a->N, a->K, ep, s)      \
                     : launch_gemm<160, EPI, 1>(tmA, tmB, a->M, a->N, a->K, ep, s);
  if (BN == 160 || BN == 192) {
    switch (a->epilogue) {
      VGPA_GEMM_DISPATCH_ODD(VGPA_EPI_BIAS)
      VGPA_GEMM_DISPATCH_ODD(VGPA_EPI_BIAS_GELU)
      VGPA_GEMM_DISPATCH_ODD(VGPA_EPI_GATE_RES)
      default:
        break;
    }
    set_error("vgpa_linear_bf16: tile width %d is not instantiated for epilogue %d", BN, a->epilogue);
    return 1;
  }
  switch (a->epilogue) {
    VGPA_GEMM_DISPATCH(VGPA_EPI_BIAS)
    VGPA_GEMM_DISPATCH(VGPA_EPI_BIAS_GELU)
    VGPA_GEMM_DISPATCH(VGPA_EPI_GATE_RES)
    VGPA_GEMM_DISPATCH(VGPA_EPI_QKV)
    VGPA_GEMM_DISPATCH(VGPA_EPI_ACCUM)
    VGPA_GEMM_DISPATCH(VGPA_EPI_GATE_RES_F32)
    default:
      break;
  }
#undef VGPA_GEMM_DISPATCH
#undef VGPA_GEMM_DISPATCH_ODD
  set_error("vgpa_linear_bf16: unknown epilogue %d", a->epilogue);
  return 1;
}
