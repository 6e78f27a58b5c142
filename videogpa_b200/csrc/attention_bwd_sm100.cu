// K2b: attention backward for head_dim 64 on tcgen05 — SURVEY.md §8 row f-2 (the DPO training step differentiates
// through F.scaled_dot_product_attention inside CogVideoXAttnProcessor2_0; train/CogVideoX-5B/03_train.py:134-157).
//
// Recompute form (no S x S tensor in HBM): the forward saves the log2-domain logsumexp L of every row; with
// delta = rowsum(dO * O),  P = exp2(scale_log2 * q k^T - L),  dS = P * (dP - delta) * scale,  dP = dO v^T:
//     dQ = dS k        dK = dS^T q        dV = P^T dO
// Two kernels, so that no output needs atomics (results are deterministic):
//   * attn_bwd_dq_kernel:  one CTA per (128-query tile, head, sample), loops over kv tiles.
//       S = Q K^T and dP = dO V^T (SS MMAs, N = 128) -> TMEM; 128 threads (one per query row) turn them into dS (bf16,
//       back into TMEM); dQ += dS K_j is a TS MMA whose B operand is the SAME K tile read MN-major.
//   * attn_bwd_dkv_kernel: one CTA per (128-key tile, head, sample), loops over query tiles, transposed problem:
//       S^T = K Q^T, dP^T = V dO^T -> TMEM; threads own key rows, L / delta of the 128 queries come from shared memory;
//       dV += P^T dO_i and dK += dS^T Q_i are TS MMAs with the dO / Q tiles read MN-major.
// TMEM: 3 buffers of 128 columns (S | dP, the bf16 dS / P^T / dS^T written in place over consumed columns) + dQ 64, or dV 64 + dK 64.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (elect.sync), warps 2-13 compute: three warps per TMEM lane quarter,
// i.e. three compute warpgroups. The streamed tiles are 64 rows and three are in flight: TMEM buffer b and warpgroup b serve
// the tiles j = b (mod 3), so the MMAs of the next tiles overlap the dS arithmetic of tile j (packed f32x2 FMA-pipe ops; the softmax
// scale is applied once in the dQ / dK epilogue).
#include "sm100.cuh"
#include "attn_common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {

constexpr int AB_NBUF = 3;                                  // TMEM buffers = compute warpgroups = streamed tiles in flight
constexpr int AB_THREADS = 64 + AB_NBUF * 128;              // TMA warp, MMA warp, AB_NBUF x 4 compute warps
constexpr int AB_T = 128;                                   // tile rows (queries or keys)
constexpr int AB_D = 64;
constexpr uint32_t AB_TILE = AB_T * AB_D * 2;               // 16384 bytes

struct BwdParams {
  const float* lse;        // [B, H, Sq]
  const float* delta;      // [B, H, Sq]
  __nv_bfloat16* dq;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  long long dq_row_stride, dq_batch_stride, dk_row_stride, dk_batch_stride, dv_row_stride, dv_batch_stride;
  int Sq, Skv, H;
  float scale, scale_log2;
};

// columns [16*c_begin, 16*c_end) of a 64-column fp32 accumulator row -> bf16 global, times `mul`
__device__ __forceinline__ void store_cols(__nv_bfloat16* dst, uint32_t tmem_addr, int c_begin, int c_end, float mul) {
  for (int c = c_begin; c < c_end; ++c) {
    uint32_t o[16];
    ptx::tmem_ld_32x16(tmem_addr + c * 16, o);
    ptx::tmem_ld_wait();
    if (dst != nullptr) {
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(__uint_as_float(o[2 * i]) * mul, __uint_as_float(o[2 * i + 1]) * mul);
      reinterpret_cast<uint4*>(dst + c * 16)[0] = make_uint4(w[0], w[1], w[2], w[3]);
      reinterpret_cast<uint4*>(dst + c * 16)[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
}

using attn::f2_pack;
using attn::f2_unpack;
using attn::f2_fma;
using attn::f2_add;
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ------------------------------------------------------------------------------------------------ dQ
// kv tiles of AB_N = 64 rows, two in flight: TMEM buffer b (S 64 + dP 64 columns) and compute warpgroup b serve the tiles
// j = b (mod 2), so the MMAs of tile j+1 run while warpgroup (j & 1) forms dS(j).
constexpr int AB_N = 64;                                    // streamed tile rows
constexpr uint32_t AB_TILE_N = AB_N * AB_D * 2;             // 8192 bytes
constexpr int AB_SLOTS = 6;
constexpr uint32_t AB_SMEM_BYTES = 2 * AB_TILE + AB_SLOTS * 2 * AB_TILE_N + AB_NBUF * 2 * AB_N * 4 * 2 + 1024 + 256;
constexpr uint32_t DQ_BUF = 128, DQ_COL_DP = 64, DQ_COL_DQ = AB_NBUF * DQ_BUF, DQ_TMEM_COLS = 512;
//   buffer b at b*128: S_b [0, 64), dP_b [64, 128); dS_b (bf16 pairs, 32 columns) overwrites S_b in place: a thread
//   stores the 16 columns of chunk c only after it has loaded S columns [32c, 32c+32), which cover them.   dQ [384, +64)

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, BwdParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + AB_TILE;
  uint8_t* sKV = sdO + AB_TILE;                         // slot s: K at s * 2 small tiles, V right after
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + AB_SLOTS * 2 * AB_TILE_N);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                         // [AB_SLOTS]
  uint64_t* kv_empty = kv_full + AB_SLOTS;              // [AB_SLOTS]
  uint64_t* sdp_full = kv_empty + AB_SLOTS;             // [AB_NBUF]
  uint64_t* ds_ready = sdp_full + AB_NBUF;              // [AB_NBUF]
  uint64_t* dq_done = ds_ready + AB_NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int m0 = blockIdx.x * AB_T;
  const int nkv = (prm.Skv + AB_N - 1) / AB_N;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmdO);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < AB_SLOTS; ++i) { ptx::mbar_init(&kv_full[i], 1); ptx::mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < AB_NBUF; ++i) { ptx::mbar_init(&sdp_full[i], 1); ptx::mbar_init(&ds_ready[i], 128); }
    ptx::mbar_init(dq_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, DQ_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, 2 * AB_TILE);
      ptx::tma_load_3d(sQ, &tmQ, q_full, head * AB_D, m0, batch);
      ptx::tma_load_3d(sdO, &tmdO, q_full, head * AB_D, m0, batch);
      for (int j = 0; j < nkv; ++j) {
        const int slot = j % AB_SLOTS;
        ptx::mbar_wait(&kv_empty[slot], ((j / AB_SLOTS) & 1) ^ 1);
        ptx::mbar_expect_tx(&kv_full[slot], 2 * AB_TILE_N);
        ptx::tma_load_3d(sKV + slot * 2 * AB_TILE_N, &tmK, &kv_full[slot], head * AB_D, j * AB_N, batch);
        ptx::tma_load_3d(sKV + slot * 2 * AB_TILE_N + AB_TILE_N, &tmV, &kv_full[slot], head * AB_D, j * AB_N, batch);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = ptx::idesc_bf16(AB_T, AB_N, 0, 0);     // [128 x 64 d] x [64 kv x 64 d]^T, both K-major
    constexpr uint32_t idesc_o = ptx::idesc_bf16(AB_T, AB_D, 0, 1);     // TMEM [128 x 64 kv] x K tile [64 kv x 64 d] read MN-major
    const uint32_t sQ_a = ptx::smem_u32(sQ), sdO_a = ptx::smem_u32(sdO), sKV_a = ptx::smem_u32(sKV);
    if (ptx::elect_one()) {
      auto issue_sdp = [&](int j) {
        const int slot = j % AB_SLOTS, b = j % AB_NBUF;
        const uint32_t sK_a = sKV_a + slot * 2 * AB_TILE_N, sV_a = sK_a + AB_TILE_N;
        ptx::mbar_wait(&kv_full[slot], (j / AB_SLOTS) & 1);
        ptx::tc_fence_after();
        {
          const uint64_t a = ptx::smem_desc_sw128(sQ_a, 16, 1024), bd = ptx::smem_desc_sw128(sK_a, 16, 1024);
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k) ptx::umma_ss(tmem_base + b * DQ_BUF, a + 2 * k, bd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        }
        {
          const uint64_t a = ptx::smem_desc_sw128(sdO_a, 16, 1024), bd = ptx::smem_desc_sw128(sV_a, 16, 1024);
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k) ptx::umma_ss(tmem_base + b * DQ_BUF + DQ_COL_DP, a + 2 * k, bd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&sdp_full[b]);
      };
      ptx::mbar_wait(q_full, 0);
      for (int j = 0; j < AB_NBUF && j < nkv; ++j) issue_sdp(j);
      for (int j = 0; j < nkv; ++j) {
        const int slot = j % AB_SLOTS, b = j % AB_NBUF;
        const uint32_t sK_a = sKV_a + slot * 2 * AB_TILE_N;
        ptx::mbar_wait(&ds_ready[b], (j / AB_NBUF) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < AB_N / 16; ++kk) {
          const uint64_t bd = ptx::smem_desc_sw128(sK_a + kk * 2048, 1024, 1024);
          ptx::umma_ts(tmem_base + DQ_COL_DQ, tmem_base + b * DQ_BUF + kk * 8, bd, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&kv_empty[slot]);
        // buffer b (S / dP, and dS in place of S) is rewritten by the MMAs of tile j + AB_NBUF, which the in-order tensor
        // pipe runs after the dQ MMAs above have read dS_b
        if (j + AB_NBUF < nkv) issue_sdp(j + AB_NBUF);
      }
      ptx::umma_commit(dq_done);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int b = (warp - 2) >> 2;                            // warpgroup = TMEM buffer; serves the kv tiles j = b (mod AB_NBUF)
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const int row = m0 + r;
    const bool live = row < prm.Sq;
    const long long stat = (static_cast<long long>(batch) * prm.H + head) * prm.Sq + (live ? row : 0);
    const float L = live ? prm.lse[stat] : INFINITY;          // rows past the end get P = 0
    const float dl = live ? prm.delta[stat] : 0.f;
    const uint64_t sc2 = f2_pack(prm.scale_log2, prm.scale_log2), negL2 = f2_pack(-L, -L), negdl2 = f2_pack(-dl, -dl);
    const int tail = prm.Skv - (nkv - 1) * AB_N;
    const uint32_t tbuf = tmem_base + lane_addr + b * DQ_BUF;
    for (int j = b; j < nkv; j += AB_NBUF) {
      ptx::mbar_wait(&sdp_full[b], (j / AB_NBUF) & 1);
      ptx::tc_fence_after();
      const int valid = (j == nkv - 1) ? tail : AB_N;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t s[32], dp[32], pk[16];
        ptx::tmem_ld_32x32(tbuf + c * 32, s);
        ptx::tmem_ld_32x32(tbuf + DQ_COL_DP + c * 32, dp);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x0, x1, d0, d1;
          f2_unpack(f2_fma(f2_pack(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, negL2), x0, x1);
          const uint64_t p2 = f2_pack(ptx::ex2_approx(x0), ptx::ex2_approx(x1));
          f2_unpack(f2_mul(p2, f2_add(f2_pack(__uint_as_float(dp[2 * i]), __uint_as_float(dp[2 * i + 1])), negdl2)), d0, d1);
          pk[i] = pack_bf16x2(d0, d1);
        }
        if (valid < AB_N) {                                   // last kv tile: columns past the end contribute nothing
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = c * 32 + 2 * i;
            if (col >= valid) pk[i] = 0u;
            else if (col + 1 >= valid) pk[i] &= 0x0000ffffu;
          }
        }
        ptx::tmem_st_32x16(tbuf + c * 16, pk);                // dS over the S columns this thread has already consumed
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ds_ready[b]);
    }
    ptx::mbar_wait(dq_done, 0);
    ptx::tc_fence_after();
    __nv_bfloat16* dst = live ? prm.dq + static_cast<long long>(batch) * prm.dq_batch_stride +
                                    static_cast<long long>(row) * prm.dq_row_stride + head * AB_D
                              : nullptr;
    // dS was formed without the scale; the 4 column chunks are split over the warpgroups (0: chunks 0-1, 1: chunk 2, 2: chunk 3)
    store_cols(dst, tmem_base + lane_addr + DQ_COL_DQ, b == 0 ? 0 : b + 1, b == 0 ? 2 : b + 2, prm.scale);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, DQ_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ dK, dV
// query tiles of 64 rows, two in flight (same ping-pong): per buffer S^T 64 + dP^T 64 + P^T 32 + dS^T 32 columns.
constexpr uint32_t KV_BUF = 128, KV_COL_DPT = 64, KV_COL_DV = AB_NBUF * KV_BUF, KV_COL_DK = KV_COL_DV + 64, KV_TMEM_COLS = 512;
//   buffer b at b*128: S^T_b [0, 64), dP^T_b [64, 128); P^T_b (bf16 pairs) overwrites S^T_b and dS^T_b overwrites dP^T_b in
//   place, 8 columns per 16-column chunk already loaded by the same thread.   dV [384, +64)   dK [448, +64)

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, BwdParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + AB_TILE;
  uint8_t* sQdO = sV + AB_TILE;                         // slot s: Q at s * 2 small tiles, dO right after
  float* sL = reinterpret_cast<float*>(sQdO + AB_SLOTS * 2 * AB_TILE_N);     // [warpgroup][2][64]: -L
  float* sDl = sL + AB_NBUF * 2 * AB_N;                                       // [warpgroup][2][64]: -delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDl + AB_NBUF * 2 * AB_N);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;                          // [AB_SLOTS]
  uint64_t* q_empty = q_full + AB_SLOTS;                // [AB_SLOTS]
  uint64_t* sdp_full = q_empty + AB_SLOTS;              // [AB_NBUF]
  uint64_t* ds_ready = sdp_full + AB_NBUF;              // [AB_NBUF]
  uint64_t* acc_done = ds_ready + AB_NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int n0 = blockIdx.x * AB_T;
  const int nq = (prm.Sq + AB_N - 1) / AB_N;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmdO);
    ptx::mbar_init(kv_full, 1);
    for (int i = 0; i < AB_SLOTS; ++i) { ptx::mbar_init(&q_full[i], 1); ptx::mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < AB_NBUF; ++i) { ptx::mbar_init(&sdp_full[i], 1); ptx::mbar_init(&ds_ready[i], 128); }
    ptx::mbar_init(acc_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, KV_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(kv_full, 2 * AB_TILE);
      ptx::tma_load_3d(sK, &tmK, kv_full, head * AB_D, n0, batch);
      ptx::tma_load_3d(sV, &tmV, kv_full, head * AB_D, n0, batch);
      for (int i = 0; i < nq; ++i) {
        const int slot = i % AB_SLOTS;
        ptx::mbar_wait(&q_empty[slot], ((i / AB_SLOTS) & 1) ^ 1);
        ptx::mbar_expect_tx(&q_full[slot], 2 * AB_TILE_N);
        ptx::tma_load_3d(sQdO + slot * 2 * AB_TILE_N, &tmQ, &q_full[slot], head * AB_D, i * AB_N, batch);
        ptx::tma_load_3d(sQdO + slot * 2 * AB_TILE_N + AB_TILE_N, &tmdO, &q_full[slot], head * AB_D, i * AB_N, batch);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = ptx::idesc_bf16(AB_T, AB_N, 0, 0);     // K / V tile [128 x 64 d] x Q / dO tile [64 q x 64 d]^T
    constexpr uint32_t idesc_o = ptx::idesc_bf16(AB_T, AB_D, 0, 1);     // TMEM [128 x 64 q] x dO / Q tile [64 q x 64 d] read MN-major
    const uint32_t sK_a = ptx::smem_u32(sK), sV_a = ptx::smem_u32(sV), sQdO_a = ptx::smem_u32(sQdO);
    if (ptx::elect_one()) {
      auto issue_sdp = [&](int i) {
        const int slot = i % AB_SLOTS, b = i % AB_NBUF;
        const uint32_t sQ_a = sQdO_a + slot * 2 * AB_TILE_N, sdO_a = sQ_a + AB_TILE_N;
        ptx::mbar_wait(&q_full[slot], (i / AB_SLOTS) & 1);
        ptx::tc_fence_after();
        {
          const uint64_t a = ptx::smem_desc_sw128(sK_a, 16, 1024), bd = ptx::smem_desc_sw128(sQ_a, 16, 1024);
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k) ptx::umma_ss(tmem_base + b * KV_BUF, a + 2 * k, bd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        }
        {
          const uint64_t a = ptx::smem_desc_sw128(sV_a, 16, 1024), bd = ptx::smem_desc_sw128(sdO_a, 16, 1024);
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k) ptx::umma_ss(tmem_base + b * KV_BUF + KV_COL_DPT, a + 2 * k, bd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&sdp_full[b]);
      };
      ptx::mbar_wait(kv_full, 0);
      for (int i = 0; i < AB_NBUF && i < nq; ++i) issue_sdp(i);
      for (int i = 0; i < nq; ++i) {
        const int slot = i % AB_SLOTS, b = i % AB_NBUF;
        const uint32_t sQ_a = sQdO_a + slot * 2 * AB_TILE_N, sdO_a = sQ_a + AB_TILE_N;
        ptx::mbar_wait(&ds_ready[b], (i / AB_NBUF) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < AB_N / 16; ++kk) {
          const uint64_t bd = ptx::smem_desc_sw128(sdO_a + kk * 2048, 1024, 1024);
          ptx::umma_ts(tmem_base + KV_COL_DV, tmem_base + b * KV_BUF + kk * 8, bd, idesc_o, (i > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < AB_N / 16; ++kk) {
          const uint64_t bd = ptx::smem_desc_sw128(sQ_a + kk * 2048, 1024, 1024);
          ptx::umma_ts(tmem_base + KV_COL_DK, tmem_base + b * KV_BUF + KV_COL_DPT + kk * 8, bd, idesc_o, (i > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&q_empty[slot]);
        if (i + AB_NBUF < nq) issue_sdp(i + AB_NBUF);
      }
      ptx::umma_commit(acc_done);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int b = (warp - 2) >> 2;                            // warpgroup = TMEM buffer; serves the query tiles i = b (mod AB_NBUF)
    const int r = quarter * 32 + lane;                        // key row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const long long stat0 = (static_cast<long long>(batch) * prm.H + head) * prm.Sq;
    const uint64_t sc2 = f2_pack(prm.scale_log2, prm.scale_log2);
    const uint32_t tbuf = tmem_base + lane_addr + b * KV_BUF;
    for (int i = b, n = 0; i < nq; i += AB_NBUF, ++n) {
      float* Lq = sL + (b * 2 + (n & 1)) * AB_N;
      float* Dq = sDl + (b * 2 + (n & 1)) * AB_N;
      if (r < AB_N) {                                         // -L and -delta of the tile's 64 queries
        const int q = i * AB_N + r;
        const bool ok = q < prm.Sq;
        Lq[r] = ok ? -prm.lse[stat0 + q] : -INFINITY;         // queries past the end: P = 0
        Dq[r] = ok ? -prm.delta[stat0 + q] : 0.f;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(b + 1) : "memory");
      ptx::mbar_wait(&sdp_full[b], n & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {                           // 16 query columns at a time
        uint32_t s[16], dp[16], pp[8], pd[8];
        ptx::tmem_ld_32x16(tbuf + c * 16, s);
        ptx::tmem_ld_32x16(tbuf + KV_COL_DPT + c * 16, dp);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int q0 = c * 16 + 2 * k;
          const uint64_t negL2 = *reinterpret_cast<const uint64_t*>(Lq + q0);
          const uint64_t negdl2 = *reinterpret_cast<const uint64_t*>(Dq + q0);
          float x0, x1, d0, d1;
          f2_unpack(f2_fma(f2_pack(__uint_as_float(s[2 * k]), __uint_as_float(s[2 * k + 1])), sc2, negL2), x0, x1);
          const float p0 = ptx::ex2_approx(x0), p1 = ptx::ex2_approx(x1);
          f2_unpack(f2_mul(f2_pack(p0, p1), f2_add(f2_pack(__uint_as_float(dp[2 * k]), __uint_as_float(dp[2 * k + 1])), negdl2)), d0, d1);
          pp[k] = pack_bf16x2(p0, p1);
          pd[k] = pack_bf16x2(d0, d1);
        }
        ptx::tmem_st_32x8(tbuf + c * 8, pp);                  // P^T over consumed S^T columns, dS^T over consumed dP^T columns
        ptx::tmem_st_32x8(tbuf + KV_COL_DPT + c * 8, pd);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ds_ready[b]);
    }
    ptx::mbar_wait(acc_done, 0);
    ptx::tc_fence_after();
    const int row = n0 + r;
    const bool live = row < prm.Skv;
    __nv_bfloat16* dvp = live ? prm.dv + static_cast<long long>(batch) * prm.dv_batch_stride +
                                    static_cast<long long>(row) * prm.dv_row_stride + head * AB_D : nullptr;
    __nv_bfloat16* dkp = live ? prm.dk + static_cast<long long>(batch) * prm.dk_batch_stride +
                                    static_cast<long long>(row) * prm.dk_row_stride + head * AB_D : nullptr;
    const int cb = b == 0 ? 0 : b + 1, ce = b == 0 ? 2 : b + 2;   // 4 column chunks over the warpgroups: 0-1, 2, 3
    store_cols(dvp, tmem_base + lane_addr + KV_COL_DV, cb, ce, 1.0f);
    store_cols(dkp, tmem_base + lane_addr + KV_COL_DK, cb, ce, prm.scale);               // dS^T was formed without the scale
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, KV_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ delta = rowsum(dO * O)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta,
                  int B, int H, int S, long long o_row, long long o_batch, long long do_row, long long do_batch) {
  const long long total = static_cast<long long>(B) * H * S;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(i % S);
    const int h = static_cast<int>((i / S) % H);
    const int b = static_cast<int>(i / (static_cast<long long>(S) * H));
    const uint4* po = reinterpret_cast<const uint4*>(o + b * o_batch + s * o_row + h * AB_D);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + b * do_batch + s * do_row + h * AB_D);
    float acc = 0.f;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const uint4 a = __ldg(po + v), g = __ldg(pd + v);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 af = unpack_bf16x2(aw[k]), gf = unpack_bf16x2(gw[k]);
        acc = fmaf(af.x, gf.x, acc);
        acc = fmaf(af.y, gf.y, acc);
      }
    }
    delta[i] = acc;                                       // layout [B, H, S]
  }
}

}  // namespace
}  // namespace vgpa

extern "C" size_t vgpa_attention_bwd_workspace_bytes(int B, int H, int Sq) {
  return static_cast<size_t>(B) * H * Sq * sizeof(float);
}

extern "C" int vgpa_attention_bwd_bf16(const vgpa_attention_bwd_args* a, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr, "vgpa_attention_bwd_bf16: null args");
  VGPA_CHECK(a->head_dim == 64, "vgpa_attention_bwd_bf16: head_dim must be 64 (got %d)", a->head_dim);
  VGPA_CHECK(a->B > 0 && a->H > 0 && a->Sq > 0 && a->Skv > 0, "vgpa_attention_bwd_bf16: bad shape B=%d H=%d Sq=%d Skv=%d", a->B, a->H, a->Sq, a->Skv);
  VGPA_CHECK(a->q && a->k && a->v && a->out && a->d_out && a->lse && a->dq && a->dk && a->dv, "vgpa_attention_bwd_bf16: null tensor pointer");
  VGPA_CHECK(a->workspace != nullptr && a->workspace_bytes >= vgpa_attention_bwd_workspace_bytes(a->B, a->H, a->Sq),
             "vgpa_attention_bwd_bf16: workspace too small (%zu bytes needed)", vgpa_attention_bwd_workspace_bytes(a->B, a->H, a->Sq));
  const int cols = a->H * 64;
  const int64_t strides[] = {a->q_row_stride, a->k_row_stride, a->v_row_stride, a->out_row_stride, a->dout_row_stride,
                             a->dq_row_stride, a->dk_row_stride, a->dv_row_stride};
  for (int64_t st : strides) VGPA_CHECK(st % 8 == 0 && st >= cols, "vgpa_attention_bwd_bf16: row strides must be multiples of 8 covering H*64 columns");
  const int64_t bstr[] = {a->q_batch_stride, a->k_batch_stride, a->v_batch_stride, a->out_batch_stride, a->dout_batch_stride,
                          a->dq_batch_stride, a->dk_batch_stride, a->dv_batch_stride};
  for (int64_t st : bstr) VGPA_CHECK(st % 8 == 0, "vgpa_attention_bwd_bf16: batch strides must be multiples of 8 elements");
  VGPA_CHECK(((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
               reinterpret_cast<uintptr_t>(a->out) | reinterpret_cast<uintptr_t>(a->d_out) | reinterpret_cast<uintptr_t>(a->dq) |
               reinterpret_cast<uintptr_t>(a->dk) | reinterpret_cast<uintptr_t>(a->dv)) & 15) == 0,
             "vgpa_attention_bwd_bf16: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* delta = static_cast<float*>(a->workspace);
  {
    const long long total = static_cast<long long>(a->B) * a->H * a->Sq;
    long long grid = (total + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    attn_delta_kernel<<<static_cast<unsigned>(grid), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(a->out), static_cast<const __nv_bfloat16*>(a->d_out), delta, a->B, a->H, a->Sq,
        a->out_row_stride, a->out_batch_stride, a->dout_row_stride, a->dout_batch_stride);
    VGPA_LAUNCH_CHECK("attn_delta_kernel");
  }
  CUtensorMap tq, tk, tv, tdo, tq_n, tk_n, tv_n, tdo_n;      // 128-row boxes (resident tiles) and 64-row boxes (streamed tiles)
  auto mk = [&](CUtensorMap* tm, const void* p, int S, int64_t row, int64_t batch, uint32_t box_rows) {
    const uint32_t box[3] = {64, box_rows, 1};
    const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)S, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)row * 2, (uint64_t)batch * 2};
    return make_tmap_bf16(tm, p, 3, dims, str, box);
  };
  if (int rc = mk(&tq, a->q, a->Sq, a->q_row_stride, a->q_batch_stride, 128)) return rc;
  if (int rc = mk(&tk, a->k, a->Skv, a->k_row_stride, a->k_batch_stride, 128)) return rc;
  if (int rc = mk(&tv, a->v, a->Skv, a->v_row_stride, a->v_batch_stride, 128)) return rc;
  if (int rc = mk(&tdo, a->d_out, a->Sq, a->dout_row_stride, a->dout_batch_stride, 128)) return rc;
  if (int rc = mk(&tq_n, a->q, a->Sq, a->q_row_stride, a->q_batch_stride, 64)) return rc;
  if (int rc = mk(&tk_n, a->k, a->Skv, a->k_row_stride, a->k_batch_stride, 64)) return rc;
  if (int rc = mk(&tv_n, a->v, a->Skv, a->v_row_stride, a->v_batch_stride, 64)) return rc;
  if (int rc = mk(&tdo_n, a->d_out, a->Sq, a->dout_row_stride, a->dout_batch_stride, 64)) return rc;
  BwdParams prm;
  prm.lse = a->lse; prm.delta = delta;
  prm.dq = static_cast<__nv_bfloat16*>(a->dq); prm.dk = static_cast<__nv_bfloat16*>(a->dk); prm.dv = static_cast<__nv_bfloat16*>(a->dv);
  prm.dq_row_stride = a->dq_row_stride; prm.dq_batch_stride = a->dq_batch_stride;
  prm.dk_row_stride = a->dk_row_stride; prm.dk_batch_stride = a->dk_batch_stride;
  prm.dv_row_stride = a->dv_row_stride; prm.dv_batch_stride = a->dv_batch_stride;
  prm.Sq = a->Sq; prm.Skv = a->Skv; prm.H = a->H;
  prm.scale = a->scale > 0.f ? a->scale : 0.125f;
  prm.scale_log2 = prm.scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    VGPA_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES));
    VGPA_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES));
    attr_set = true;
  }
  attn_bwd_dq_kernel<<<dim3((a->Sq + AB_T - 1) / AB_T, a->H, a->B), AB_THREADS, AB_SMEM_BYTES, s>>>(tq, tk_n, tv_n, tdo, prm);
  VGPA_LAUNCH_CHECK("attn_bwd_dq_kernel");
  attn_bwd_dkv_kernel<<<dim3((a->Skv + AB_T - 1) / AB_T, a->H, a->B), AB_THREADS, AB_SMEM_BYTES, s>>>(tq_n, tk, tv, tdo_n, prm);
  VGPA_LAUNCH_CHECK("attn_bwd_dkv_kernel");
  return 0;
}
