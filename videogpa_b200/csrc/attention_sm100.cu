// K1: joint text+video full attention, head_dim 64, for the CogVideoX DiT block.
//
// Replaces F.scaled_dot_product_attention inside diffusers' CogVideoXAttnProcessor2_0
// (reached from generate/CogVideoX-5B.py:72-77 and
// train/CogVideoX-5B/03_train.py:134-151). q/k arrive already LayerNorm'ed and rotated by the
// fused QKV GEMM epilogue (gemm_sm100.cu), so this kernel is softmax(q k^T / sqrt(d)) v only.
//
// One CTA = 256 query rows of one (batch, head): two 128-row Q tiles. 384 threads:
//   warpgroup 0: warp 0 = TMA producer (Q once, then K / V tiles through a 6-slot ring of 16 KB tiles),
//                warp 1 = tcgen05 issuer under elect.sync (S_t = Q_t K_j^T, O_t += P_t V_j, all in TMEM),
//                warps 2-3 idle (they only donate registers)
//   warpgroup 1: softmax of Q tile 0 (one thread = one query row, straight out of its TMEM lane)
//   warpgroup 2: softmax of Q tile 1
// TMEM columns (all 512 used): S_t [t*128, +128) fp32;  P_t half hh [256 + t*64 + hh*32, +32) bf16 pairs;
// O_t [384 + t*64, +64) fp32. P never touches shared memory: the softmax threads store it back to TMEM
// (tcgen05.st) and the PV product is a TS-form tcgen05.mma (A operand from TMEM), which halves the shared-memory
// traffic of the kernel.
//
// With head_dim 64 there are only 128 MMA FLOPs per softmax element, so the softmax warpgroups (MUFU ex2 at
// 16/clk/SM, FMA/ALU issue) are the bound, not the tensor pipe. The kernel is built to keep them busy:
//   * split-half software pipelining: a thread holds the row as two 64-column register halves; while it works on
//     one half the other half of this tile / the first half of the next tile streams in from TMEM, and the S tile
//     is released (s_free) as soon as the row is in registers, so S_t(j+1) is computed under softmax(j);
//   * P is published per half (p_ready[t][hh]) and the wait::st + arrive is deferred into the next half's
//     instruction stream; PV runs per half (K = 64);
//   * lazy rescale: the exponent offset may go stale by up to 2^8 (O / l are rescaled only when a half-row max
//     grows by more than that), which keeps the TMEM read-modify-write of O off the steady-state path; the final
//     O / l is exact;
//   * scale/subtract, row-sum and (part of) the exponentials run as packed f32x2 FMA-pipe ops; NPOLY of the 64
//     column pairs of a row use a Cody-Waite + degree-3 polynomial exp2 on the FMA pipe instead of MUFU (relative
//     error ~1e-4, below the bf16 rounding of P); the row max uses 3-input FMNMX3;
//   * the issuer walks the events of the two Q tiles in an order that keeps warpgroup 2 about half a tile behind
//     warpgroup 1, so the MUFU-idle windows of the two warps sharing a scheduler do not coincide.
// Measurements and the design history: profiles/r01_attn_ncu_summary.md.
#include "sm100.cuh"
#include "attn_common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>
#include <stdlib.h>

namespace vgpa {
namespace {

constexpr int AT_THREADS = 384;
constexpr int AT_BM = 128;        // query rows per tile
constexpr int AT_BN = 128;        // kv rows per tile
constexpr int AT_D = 64;          // head dim
constexpr int AT_KV_SLOTS = 6;    // ring of 16 KB tiles: K0 V0 K1 V1 ...
constexpr uint32_t AT_TILE_BYTES = AT_BN * AT_D * 2;        // 16384
constexpr uint32_t AT_SMEM_BYTES = 2 * AT_TILE_BYTES + AT_KV_SLOTS * AT_TILE_BYTES + 1024 + 256;
constexpr uint32_t AT_TMEM_COLS = 512;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;  // log2 units

struct AttnParams {
  __nv_bfloat16* out;
  long long out_row_stride;
  long long out_batch_stride;
  int Sq, Skv;
  float scale_log2;
  float* lse;          // optional [B, H, Sq]: log2-domain logsumexp of the scaled scores (for the backward kernels)
  int H;
  const float* bounds; // non-null: the bounded-softmax kernel runs too; skip the (batch, head)s it serves
};

using namespace attn;

// exp2 of NPAIRS column pairs starting at pair PAIR0 of the row (for the poly pattern) held in s[2*i], s[2*i+1];
// packs P as bf16x2 into pk[i] and accumulates the row sum into l2a/l2b.
template <int NPOLY, int PAIR0, int NPAIRS>
__device__ __forceinline__ void exp_pairs(const uint32_t* s, uint32_t* pk, uint64_t sc2, uint64_t negm2,
                                          uint64_t& l2a, uint64_t& l2b) {
#pragma unroll
  for (int i = 0; i < NPAIRS; ++i) {
    const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, negm2);
    float p0, p1;
    if (pair_uses_poly<NPOLY>(PAIR0 + i)) {
      ex2_poly2(x2, p0, p1);
    } else {
      float x0, x1;
      f2_unpack(x2, x0, x1);
      p0 = ptx::ex2_approx(x0);
      p1 = ptx::ex2_approx(x1);
    }
    if (i & 1) l2b = f2_add(l2b, f2_pack(p0, p1)); else l2a = f2_add(l2a, f2_pack(p0, p1));
    pk[i] = pack_bf16x2(p0, p1);
  }
}

__device__ __forceinline__ float max64(const uint32_t (&s)[64]) {
  float mx0 = max3(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
  float mx1 = max3(__uint_as_float(s[3]), __uint_as_float(s[4]), __uint_as_float(s[5]));
  float mx2 = max3(__uint_as_float(s[6]), __uint_as_float(s[7]), __uint_as_float(s[8]));
  float mx3 = max3(__uint_as_float(s[9]), __uint_as_float(s[10]), __uint_as_float(s[11]));
#pragma unroll
  for (int i = 12; i < 60; i += 8) {
    mx0 = max3(mx0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
    mx1 = max3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
    mx2 = max3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
    mx3 = max3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
  }
  mx0 = max3(mx0, __uint_as_float(s[60]), __uint_as_float(s[61]));
  mx1 = max3(mx1, __uint_as_float(s[62]), __uint_as_float(s[63]));
  return fmaxf(max3(mx0, mx1, mx2), mx3);
}

// TMEM columns: S_t [t*128, t*128+128)   P_t half hh [256 + t*64 + hh*32, +32) (bf16 pairs)   O_t [384 + t*64, +64)
constexpr uint32_t AT_COL_P = 256;
constexpr uint32_t AT_COL_O = 384;

template <int NPOLY>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_d64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, AttnParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + 2 * AT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + AT_KV_SLOTS * AT_TILE_BYTES);
  uint64_t* q_full = bars;                        // 1
  uint64_t* kv_full = bars + 1;                   // AT_KV_SLOTS
  uint64_t* kv_empty = kv_full + AT_KV_SLOTS;     // AT_KV_SLOTS
  uint64_t* s_full = kv_empty + AT_KV_SLOTS;      // [2]    S_t(j) is in TMEM
  uint64_t* s_free = s_full + 2;                  // [2]    every softmax thread of tile t holds its S row in registers
  uint64_t* p_ready = s_free + 2;                 // [2][2] P_t(j, half) is in TMEM
  uint64_t* pv_done = p_ready + 4;                // [2][2] O_t += P_t(j, half) V_j(half) has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int m0 = blockIdx.x * (2 * AT_BM);
  const int nkv = (prm.Skv + AT_BN - 1) / AT_BN;
  if (prm.bounds != nullptr && bounded_m(prm.bounds, batch * prm.H + head, prm.scale_log2) <= kBoundedMax) return;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < AT_KV_SLOTS; ++i) {
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_free[i], 128);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&p_ready[i], 128);
      ptx::mbar_init(&pv_done[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, AT_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (wg == 0) {
    ptx::setmaxnreg_dec<56>();
    // Ring order of the 16 KB tiles: K_0, then for every j: K_{j+1} (if any), V_j.
    if (warp == 0) {
      // ---------------------------------------------------------- TMA producer
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(q_full, 2 * AT_TILE_BYTES);
        ptx::tma_load_3d(sQ, &tmQ, q_full, head * AT_D, m0, batch);
        ptx::tma_load_3d(sQ + AT_TILE_BYTES, &tmQ, q_full, head * AT_D, m0 + AT_BM, batch);
        int slot = 0;
        uint32_t phase = 0;
        auto load = [&](const CUtensorMap* tm, int row0) {
          ptx::mbar_wait(&kv_empty[slot], phase ^ 1);
          ptx::mbar_expect_tx(&kv_full[slot], AT_TILE_BYTES);
          ptx::tma_load_3d(sKV + slot * AT_TILE_BYTES, tm, &kv_full[slot], head * AT_D, row0, batch);
          if (++slot == AT_KV_SLOTS) { slot = 0; phase ^= 1; }
        };
        load(&tmK, 0);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) load(&tmK, (j + 1) * AT_BN);
          load(&tmV, j * AT_BN);
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- tcgen05 issuer
      constexpr uint32_t idesc_s = ptx::idesc_bf16(AT_BM, AT_BN, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = ptx::idesc_bf16(AT_BM, AT_D, 0, 1);   // P (TMEM) x V (MN-major)
      const uint32_t sQ_a = ptx::smem_u32(sQ);
      const uint32_t sKV_a = ptx::smem_u32(sKV);
      if (ptx::elect_one()) {   // elect.sync: the compiler then issues tcgen05 ops without a per-lane waterfall loop
        // position of a tile in the ring sequence K_0, K_1, V_0, K_2, V_1, ..., K_{n-1}, V_{n-2}, V_{n-1}
        auto idx_k = [&](int j) { return j == 0 ? 0 : 2 * j - 1; };
        auto idx_v = [&](int j) { return j < nkv - 1 ? 2 * j + 2 : 2 * nkv - 1; };
        auto kv_release = [&](int idx) { ptx::umma_commit(&kv_empty[idx % AT_KV_SLOTS]); };
        auto do_s = [&](int t, int idx) {
          const uint64_t a = ptx::smem_desc_sw128(sQ_a + t * AT_TILE_BYTES, 16, 1024);
          const uint64_t b = ptx::smem_desc_sw128(sKV_a + (idx % AT_KV_SLOTS) * AT_TILE_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k)
            ptx::umma_ss(tmem_base + t * AT_BN, a + 2 * k, b + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          ptx::umma_commit(&s_full[t]);
        };
        auto do_pv = [&](int t, int hh, int idx, bool first) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t b = ptx::smem_desc_sw128(sKV_a + (idx % AT_KV_SLOTS) * AT_TILE_BYTES + (hh * 4 + kk) * 2048, 1024, 1024);
            ptx::umma_ts(tmem_base + AT_COL_O + t * AT_D, tmem_base + AT_COL_P + t * 64 + hh * 32 + kk * 8, b, idesc_o,
                         (first && kk == 0) ? 0u : 1u);
          }
          ptx::umma_commit(&pv_done[t * 2 + hh]);
        };
        ptx::mbar_wait(q_full, 0);
        ptx::mbar_wait(&kv_full[0], 0);
        ptx::tc_fence_after();
        do_s(0, 0);
        do_s(1, 0);
        kv_release(0);
        {
          // Anti-phase order: warpgroup 1 is held about half a kv tile behind warpgroup 0 (its S tile is only
          // issued after warpgroup 0 has published its first P half), so the MUFU-idle windows of the two
          // warps that share an SM sub-partition (tile boundary, waits) never coincide.
          auto wait_kv = [&](int idx) { ptx::mbar_wait(&kv_full[idx % AT_KV_SLOTS], (idx / AT_KV_SLOTS) & 1); };
          auto pv = [&](int t, int hh, int j) {
            ptx::mbar_wait(&p_ready[t * 2 + hh], j & 1);
            ptx::tc_fence_after();
            do_pv(t, hh, idx_v(j), j == 0 && hh == 0);
          };
          auto sq = [&](int t, int j) {   // S_t(j+1)
            ptx::mbar_wait(&s_free[t], j & 1);
            ptx::tc_fence_after();
            do_s(t, idx_k(j + 1));
          };
          for (int j = 0; j < nkv; ++j) {
            if (j > 0) pv(0, 1, j - 1);
            if (j + 1 < nkv) { wait_kv(idx_k(j + 1)); sq(0, j); }
            wait_kv(idx_v(j));
            pv(0, 0, j);
            if (j > 0) { pv(1, 1, j - 1); kv_release(idx_v(j - 1)); }
            if (j + 1 < nkv) { sq(1, j); kv_release(idx_k(j + 1)); }
            pv(1, 0, j);
          }
          pv(0, 1, nkv - 1);
          pv(1, 1, nkv - 1);
          kv_release(idx_v(nkv - 1));
        }
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups
    ptx::setmaxnreg_inc<224>();
    const int t = wg - 1;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // row inside the Q tile
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + t * AT_BN;
    const uint32_t tP = tmem_base + lane_addr + AT_COL_P + t * 64;
    const uint32_t tO = tmem_base + lane_addr + AT_COL_O + t * AT_D;
    const float sc = prm.scale_log2;
    const uint64_t sc2 = f2_pack(sc, sc);
    const int tail = prm.Skv - (nkv - 1) * AT_BN;            // valid kv rows in the last tile
    float m_used = -INFINITY;
    uint64_t l2a = f2_pack(0.f, 0.f), l2b = l2a;             // 4 partial row sums

    uint32_t sa[64], sb[64];                                 // S row: columns [0,64) and [64,128)

    // Lazy rescale: the exponent offset m_used only moves when a half-row max exceeds it by 2^8.
    // P_t(j, hh) is stored to TMEM right after its exponentials, but the wait::st + p_ready arrive is
    // deferred into the next half's instruction stream so the store latency hides behind MUFU work.
    bool pend = false;                                        // a P store of this thread awaits publication
    int pend_hh = 0;
    auto publish_pending = [&]() {
      if (pend) {
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&p_ready[t * 2 + pend_hh]);
        pend = false;
      }
    };
    auto rescale_check = [&](float hmax, int j, int hh) {
      const bool need = hmax * sc > m_used + AT_RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        publish_pending();                                    // the PV we are about to wait for may need it
        const float m_new = fmaxf(m_used, hmax * sc);
        const float factor = ptx::ex2_approx(m_used - m_new);   // first half of all: exp2(-inf) = 0
        m_used = m_new;
        const uint64_t f2 = f2_pack(factor, factor);
        const uint64_t z2 = f2_pack(0.f, 0.f);
        l2a = f2_fma(l2a, f2, z2);
        l2b = f2_fma(l2b, f2, z2);
        if (j > 0 || hh > 0) {
          if (hh == 0) ptx::mbar_wait(&pv_done[t * 2 + 1], (j - 1) & 1);
          else ptx::mbar_wait(&pv_done[t * 2 + 0], j & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < AT_D / 16; ++c) {
            uint32_t o[16];
            ptx::tmem_ld_32x16(tO + c * 16, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            ptx::tmem_st_32x16(tO + c * 16, o);
          }
          ptx::tmem_st_wait();
        }
      }
    };
    auto mask_tail = [&](uint32_t (&v)[64], int col0) {
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (col0 + i >= tail) v[i] = 0xff800000u;  // -inf
    };

    ptx::mbar_wait(&s_full[t], 0);
    ptx::tc_fence_after();
    ptx::tmem_ld_32x32(tS, *reinterpret_cast<uint32_t (*)[32]>(&sa[0]));
    ptx::tmem_ld_32x32(tS + 32, *reinterpret_cast<uint32_t (*)[32]>(&sa[32]));
    ptx::tmem_ld_wait();
    if (nkv == 1 && tail < 64) mask_tail(sa, 0);
    float hmax_a = max64(sa);
    uint32_t pk[32];

    for (int j = 0; j < nkv; ++j) {
      const bool last = (j == nkv - 1);
      // ================= half 0: columns [0,64) are in sa; [64,128) stream into sb meanwhile
      ptx::tmem_ld_32x32(tS + 64, *reinterpret_cast<uint32_t (*)[32]>(&sb[0]));
      ptx::tmem_ld_32x32(tS + 96, *reinterpret_cast<uint32_t (*)[32]>(&sb[32]));
      rescale_check(hmax_a, j, 0);
      float hmax_b;
      {
        const uint64_t negm2 = f2_pack(-m_used, -m_used);
        exp_pairs<NPOLY, 0, 16>(&sa[0], &pk[0], sc2, negm2, l2a, l2b);
        publish_pending();                                   // P_t(j-1, 1)
        ptx::tmem_ld_wait();                                 // sb has landed: S_t(j) is fully in registers
        ptx::tc_fence_before();
        ptx::mbar_arrive(&s_free[t]);                        // S_t(j+1) may now overwrite the TMEM tile
        if (last && tail < AT_BN) mask_tail(sb, 64);
        hmax_b = max64(sb);
        exp_pairs<NPOLY, 16, 16>(&sa[32], &pk[16], sc2, negm2, l2a, l2b);
      }
      if (j > 0) {                                           // P_t(j-1, 0) has been consumed
        ptx::mbar_wait(&pv_done[t * 2 + 0], (j - 1) & 1);
        ptx::tc_fence_after();
      }
      ptx::tmem_st_32x32(tP, pk);
      pend = true; pend_hh = 0;
      // ================= half 1: columns [64,128) are in sb; the first half of S_t(j+1) streams into sa
      rescale_check(hmax_b, j, 1);
      {
        const uint64_t negm2 = f2_pack(-m_used, -m_used);
        exp_pairs<NPOLY, 32, 16>(&sb[0], &pk[0], sc2, negm2, l2a, l2b);
        publish_pending();                                   // P_t(j, 0)
        if (!last) {
          ptx::mbar_wait(&s_full[t], (j + 1) & 1);
          ptx::tc_fence_after();
          ptx::tmem_ld_32x32(tS, *reinterpret_cast<uint32_t (*)[32]>(&sa[0]));
          ptx::tmem_ld_32x32(tS + 32, *reinterpret_cast<uint32_t (*)[32]>(&sa[32]));
        }
        exp_pairs<NPOLY, 48, 16>(&sb[32], &pk[16], sc2, negm2, l2a, l2b);
      }
      if (j > 0) {
        ptx::mbar_wait(&pv_done[t * 2 + 1], (j - 1) & 1);
        ptx::tc_fence_after();
      }
      ptx::tmem_st_32x32(tP + 32, pk);
      pend = true; pend_hh = 1;
      if (!last) {
        ptx::tmem_ld_wait();                                 // sa = first half of S_t(j+1)
        if (j + 1 == nkv - 1 && tail < 64) mask_tail(sa, 0);
        hmax_a = max64(sa);
      }
    }
    publish_pending();

    // ---------------------------------------------------------- epilogue: O / l -> bf16 global
    ptx::mbar_wait(&pv_done[t * 2 + 1], (nkv - 1) & 1);
    ptx::tc_fence_after();
    float la, lb, lc, ld;
    f2_unpack(l2a, la, lb);
    f2_unpack(l2b, lc, ld);
    const float l = (la + lb) + (lc + ld);
    const int row = m0 + t * AT_BM + r;
    const float inv_l = 1.0f / l;
    if (prm.lse != nullptr && row < prm.Sq)
      prm.lse[(static_cast<long long>(batch) * prm.H + head) * prm.Sq + row] = m_used + log2f(l);
    __nv_bfloat16* orow = prm.out + static_cast<long long>(batch) * prm.out_batch_stride +
                          static_cast<long long>(row < prm.Sq ? row : 0) * prm.out_row_stride + head * AT_D;
#pragma unroll
    for (int c = 0; c < AT_D / 16; ++c) {
      uint32_t o[16];
      ptx::tmem_ld_32x16(tO + c * 16, o);
      ptx::tmem_ld_wait();
      if (row < prm.Sq) {
        uint4 v0, v1;
        v0.x = pack_bf16x2(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l);
        v0.y = pack_bf16x2(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l);
        v0.z = pack_bf16x2(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l);
        v0.w = pack_bf16x2(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l);
        v1.x = pack_bf16x2(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l);
        v1.y = pack_bf16x2(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l);
        v1.z = pack_bf16x2(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l);
        v1.w = pack_bf16x2(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l);
        reinterpret_cast<uint4*>(orow + c * 16)[0] = v0;
        reinterpret_cast<uint4*>(orow + c * 16)[1] = v1;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, AT_TMEM_COLS);
}

template <int NPOLY>
int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& prm, dim3 grid,
                cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VGPA_CUDA(cudaFuncSetAttribute(attn_fwd_d64_kernel<NPOLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES));
    attr_set = true;
  }
  attn_fwd_d64_kernel<NPOLY><<<grid, AT_THREADS, AT_SMEM_BYTES, stream>>>(tq, tk, tv, prm);
  VGPA_LAUNCH_CHECK("attn_fwd_d64_kernel");
  return 0;
}

}  // namespace
}  // namespace vgpa

namespace vgpa {
int launch_attention_d128(const vgpa_attention_args* a, cudaStream_t stream);
size_t attention_d64_workspace_bytes(int B, int H);
int launch_attention_d64_bounded(const vgpa_attention_args* a, float* bounds, float scale_log2, int npoly8, cudaStream_t stream);
}

extern "C" size_t vgpa_attention_workspace_bytes(int B, int H, int head_dim) {
  return head_dim == 64 ? vgpa::attention_d64_workspace_bytes(B, H) : 0;
}

extern "C" int vgpa_attention_bf16(const vgpa_attention_args* a, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr, "vgpa_attention_bf16: null args");
  VGPA_CHECK(a->head_dim == 64 || a->head_dim == 128, "vgpa_attention_bf16: head_dim must be 64 or 128 (got %d)", a->head_dim);
  VGPA_CHECK(a->B > 0 && a->H > 0 && a->Sq > 0 && a->Skv > 0, "vgpa_attention_bf16: bad shape B=%d H=%d Sq=%d Skv=%d",
             a->B, a->H, a->Sq, a->Skv);
  VGPA_CHECK(a->q && a->k && a->v && a->out, "vgpa_attention_bf16: null tensor pointer");
  VGPA_CHECK(a->q_row_stride % 8 == 0 && a->k_row_stride % 8 == 0 && a->v_row_stride % 8 == 0 &&
                 a->out_row_stride % 8 == 0 && a->q_batch_stride % 8 == 0 && a->k_batch_stride % 8 == 0 &&
                 a->v_batch_stride % 8 == 0 && a->out_batch_stride % 8 == 0,
             "vgpa_attention_bf16: strides must be multiples of 8 elements");
  VGPA_CHECK(((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) |
               reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->out)) & 15) == 0,
             "vgpa_attention_bf16: pointers must be 16-byte aligned");
  const int cols = a->H * a->head_dim;
  VGPA_CHECK(a->q_row_stride >= cols && a->k_row_stride >= cols && a->v_row_stride >= cols && a->out_row_stride >= cols,
             "vgpa_attention_bf16: row strides must cover H*head_dim columns");
  VGPA_CHECK(a->lse == nullptr || a->head_dim == 64, "vgpa_attention_bf16: the logsumexp output is only available for head_dim 64");
  if (a->head_dim == 128) return launch_attention_d128(a, static_cast<cudaStream_t>(stream));
  CUtensorMap tq, tk, tv;
  const uint32_t box[3] = {64, 128, 1};
  {
    const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)a->Sq, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->q_row_stride * 2, (uint64_t)a->q_batch_stride * 2};
    if (int rc = make_tmap_bf16(&tq, a->q, 3, dims, str, box)) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)a->Skv, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->k_row_stride * 2, (uint64_t)a->k_batch_stride * 2};
    if (int rc = make_tmap_bf16(&tk, a->k, 3, dims, str, box)) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)a->Skv, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->v_row_stride * 2, (uint64_t)a->v_batch_stride * 2};
    if (int rc = make_tmap_bf16(&tv, a->v, 3, dims, str, box)) return rc;
  }
  AttnParams prm;
  prm.out = static_cast<__nv_bfloat16*>(a->out);
  prm.out_row_stride = a->out_row_stride;
  prm.out_batch_stride = a->out_batch_stride;
  prm.Sq = a->Sq;
  prm.Skv = a->Skv;
  prm.lse = a->lse;
  prm.H = a->H;
  prm.bounds = nullptr;
  const float scale = a->scale > 0.f ? a->scale : 0.125f;
  prm.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((a->Sq + 2 * AT_BM - 1) / (2 * AT_BM), a->H, a->B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // Development knob: share of the exponentials evaluated on the FMA pipe, in 64ths of a row (default 16).
  static int npoly = -1;
  if (npoly < 0) {
    const char* e = getenv("VGPA_ATTN_NPOLY");
    npoly = e ? atoi(e) : 16;
  }
  // Bounded-softmax fast path (attention_d64b_sm100.cu) when the caller provides the scratch for the |q|, |k| bounds.
  // Development knobs: VGPA_ATTN_FAST=0 forces the exact kernel, VGPA_ATTN_NPOLY8 = FMA-pipe share of the exponentials in 8ths.
  static int fast = -1, npoly8 = 4;
  if (fast < 0) {
    const char* e = getenv("VGPA_ATTN_FAST");
    fast = e ? atoi(e) : 1;
    const char* e8 = getenv("VGPA_ATTN_NPOLY8");
    if (e8) npoly8 = atoi(e8);
  }
  if (fast && a->workspace != nullptr) {
    VGPA_CHECK(a->workspace_bytes >= attention_d64_workspace_bytes(a->B, a->H),
               "vgpa_attention_bf16: workspace too small (%zu bytes, need %zu)", a->workspace_bytes,
               attention_d64_workspace_bytes(a->B, a->H));
    VGPA_CHECK((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0, "vgpa_attention_bf16: workspace must be 16-byte aligned");
    float* bounds = static_cast<float*>(a->workspace);
    if (int rc = launch_attention_d64_bounded(a, bounds, prm.scale_log2, npoly8, s)) return rc;
    prm.bounds = bounds;
  }
  switch (npoly) {
    case 0: return launch_attn<0>(tq, tk, tv, prm, grid, s);
    case 32: return launch_attn<32>(tq, tk, tv, prm, grid, s);
    case 24: return launch_attn<24>(tq, tk, tv, prm, grid, s);
    default: return launch_attn<16>(tq, tk, tv, prm, grid, s);
  }
}
