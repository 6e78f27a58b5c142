// K1 (head_dim 64), bounded-softmax forward kernel: the fast path of vgpa_attention_bf16.
//
// Replaces F.scaled_dot_product_attention inside diffusers' CogVideoXAttnProcessor2_0 (reached from
// generate/CogVideoX-5B.py:72-77 and train/CogVideoX-5B/03_train.py:134-151), like attention_sm100.cu.
//
// Why a second kernel. With head_dim 64 there are 128 MMA FLOPs per softmax element, so the softmax warps, not the tensor
// pipe, bound the kernel, and in the online-softmax kernel (attention_sm100.cu) a third of their instructions maintain the
// running row maximum (FMNMX3 tree, rescale vote, lazy-rescale branch) and the row sum (FADD2) (profiles/r01_attn_ncu_summary.md).
// softmax is shift invariant, so any per-row offset m gives the same result as the row maximum as long as exp2(x - m) neither
// overflows nor underflows for the entries that matter. CogVideoX applies a per-head LayerNorm to q and k (qk_norm), which
// bounds |q.k| <= |q||k| (Cauchy-Schwarz). A pre-pass (`attn_qk_bound_kernel`) measures max|q|^2 and max|k|^2 per (batch, head);
// with M = ceil(max|q| max|k| scale log2e) and m = M - 64 every exponent lies in [-2M + 63, 65]: no overflow in fp32 / bf16
// (8-bit exponent), and the row's largest term is >= 2^(63 - 2M), so for M <= 90 nothing that contributes to the fp32 sum is
// flushed. A (batch, head) with M > 90 is left to the exact kernel (both kernels are launched; each CTA reads the bound
// and returns at once if the head belongs to the other kernel). There is no per-row state at all:
//   * no row maximum, no rescale, no dependency between the columns of a row: a thread streams 16-column chunks
//     TMEM -> exp2 -> bf16 -> TMEM;
//   * the row sum comes out of the tensor pipe: every PV k-step is followed by an N = 16 MMA of the same P columns with an
//     all-ones B tile, accumulating l in TMEM next to O (the sum of the bf16-rounded P, i.e. of what the PV product saw);
//   * m is an integer, so floor(x) for the polynomial exp2 comes from one fma.rm against (1.5*2^23 - m) and the fraction from
//     one more FMA: 6 FMA-pipe ops + 2 integer ops per column pair, no clamping (x >= -126 is guaranteed by M <= 90).
// Since no row statistics are exchanged, the kv columns of a Q tile are split over two warpgroups at no cost:
// 640 threads = 4 softmax warpgroups (Q tile t, column half hh: four softmax warps per SM sub-partition) + a control
// warpgroup placed LAST (TMA producer, one tcgen05-issuing thread per Q tile): the warp scheduler prefers the highest warp id
// among eligible warps, and an issuer that queues behind busy softmax warps delays every MMA. kv tiles are 96 rows so that
// S (2 x 96), P (4 x 24), O (2 x 64) and l (2 x 16) fit the 512 TMEM columns. Registers: 640 x 96 at launch; the control
// warpgroup drops to 32 and the softmax warpgroups rise to 112 (128 x 64 released = 512 x 16 claimed).
// History and measurements: profiles/r02_attn_ncu_summary.md.
#include "sm100.cuh"
#include "attn_common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>
#include <stdlib.h>

namespace vgpa {
namespace {
using namespace attn;

constexpr int FB_THREADS = 640;
constexpr int FB_BM = 128;
constexpr int FB_D = 64;
constexpr int FB_BN = 96;                       // kv rows per tile
constexpr int FB_HALF = FB_BN / 2;              // kv columns of a softmax warpgroup
constexpr int FB_KSTEPS = FB_HALF / 16;         // MMA k-steps of a P half
constexpr int FB_SLOTS = 8;                     // K/V ring slots
constexpr uint32_t FB_Q_BYTES = FB_BM * FB_D * 2;              // 16384
constexpr uint32_t FB_TILE_BYTES = FB_BN * FB_D * 2;           // 12288
constexpr uint32_t FB_ONES_BYTES = 2048;
constexpr uint32_t FB_SMEM_BYTES = 2 * FB_Q_BYTES + FB_SLOTS * FB_TILE_BYTES + FB_ONES_BYTES + 1024 + 256;
constexpr uint32_t FB_TMEM_COLS = 512;
constexpr float FB_OFFSET = 64.0f;              // m = M - 64

// TMEM columns. S_t: 96 fp32 columns; P_(t,hh): 24 columns of bf16 pairs; O_t: 64; l_t: 16.
__host__ __device__ constexpr uint32_t col_s(int t) { return t * 128; }
__host__ __device__ constexpr uint32_t col_p(int t, int hh) { return hh == 0 ? 96u + 128u * t : 384u + 32u * t; }
__host__ __device__ constexpr uint32_t col_o(int t) { return 256 + t * 64; }
__host__ __device__ constexpr uint32_t col_l(int t) { return 448 + t * 16; }

struct FbParams {
  __nv_bfloat16* out;
  long long out_row_stride;
  long long out_batch_stride;
  int Sq, Skv;
  float scale_log2;
  float* lse;
  int H;
  const float* bounds;   // [B*H][2]: max |q|^2, max |k|^2
};

__device__ __forceinline__ uint64_t f2_fma_rm(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// exp2(s*c - m) of the column pairs of one 16-column chunk; NP of the 8 pairs on the FMA pipe.
//   c2 = (c, c), negm2 = (-m, -m), a2 = (1.5*2^23 - m) twice.
template <int NP, bool MASK>
__device__ __forceinline__ void exp_chunk(const uint32_t (&s)[16], uint32_t* pk, uint64_t c2, uint64_t negm2, uint64_t a2,
                                          int col0, int tail) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint64_t s2 = f2_pack(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
    float p0, p1;
    const bool poly = !MASK && (((i + 1) * NP) / 8 != (i * NP) / 8);
    if (poly) {
      const uint64_t xr = f2_fma_rm(s2, c2, a2);               // 1.5*2^23 + floor(x)
      const uint64_t nt = f2_sub(a2, xr);                       // -(m + floor(x)), exact
      const uint64_t fr = f2_fma(s2, c2, nt);                   // x - floor(x) in [0, 1)
      uint64_t acc = f2_fma(fr, f2_pack(0.077119089663028717f, 0.077119089663028717f),
                            f2_pack(0.227564394474029541f, 0.227564394474029541f));
      acc = f2_fma(acc, fr, f2_pack(0.695146143436431885f, 0.695146143436431885f));
      acc = f2_fma(acc, fr, f2_pack(1.0f, 1.0f));
      float r0, r1, q0, q1;
      f2_unpack(xr, r0, r1);
      f2_unpack(acc, q0, q1);
      p0 = __int_as_float((__float_as_int(r0) << 23) + __float_as_int(q0));
      p1 = __int_as_float((__float_as_int(r1) << 23) + __float_as_int(q1));
    } else {
      float x0, x1;
      f2_unpack(f2_fma(s2, c2, negm2), x0, x1);
      if (MASK) {
        if (col0 + 2 * i >= tail) x0 = -INFINITY;
        if (col0 + 2 * i + 1 >= tail) x1 = -INFINITY;
      }
      p0 = ptx::ex2_approx(x0);
      p1 = ptx::ex2_approx(x1);
    }
    pk[i] = pack_bf16x2(p0, p1);
  }
}

template <int NP>
__global__ void __launch_bounds__(FB_THREADS, 1)
attn_fwd_d64_bounded_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, FbParams prm) {
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  // Offset of this (batch, head). A head whose bound is too large belongs to the exact kernel.
  const float mbound = bounded_m(prm.bounds, batch * prm.H + head, prm.scale_log2);
  if (!(mbound <= kBoundedMax)) return;
  const float m_off = mbound - FB_OFFSET;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + 2 * FB_Q_BYTES;
  uint8_t* sOnes = sKV + FB_SLOTS * FB_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + FB_ONES_BYTES);
  uint64_t* q_full = bars;                        // 1
  uint64_t* kv_full = bars + 1;                   // FB_SLOTS
  uint64_t* kv_empty = kv_full + FB_SLOTS;        // FB_SLOTS
  uint64_t* s_full = kv_empty + FB_SLOTS;         // [2]    S_t(j) is in TMEM
  uint64_t* s_free = s_full + 2;                  // [2]    8 warps: both column halves of tile t hold S_t(j) in registers
  uint64_t* p_ready = s_free + 2;                 // [2][2] 4 warps: P_(t,hh)(j) is in TMEM
  uint64_t* pv_done = p_ready + 4;                // [2][2] O_t += P_(t,hh)(j) V_j(hh) has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;                        // 0..3 softmax (Q tile wg >> 1, column half wg & 1), 4 = control
  const int cw = warp - 16;                        // control warpgroup: 0 = TMA producer, 1 = tcgen05 issuer
  const int m0 = blockIdx.x * (2 * FB_BM);
  const int nkv = (prm.Skv + FB_BN - 1) / FB_BN;

  if (cw == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < FB_SLOTS; ++i) {
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_free[i], 8);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&p_ready[i], 4);
      ptx::mbar_init(&pv_done[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (cw == 1) {
    ptx::tmem_alloc(tmem_slot, FB_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (warp < 4) {                                 // all-ones B tile of the row-sum MMA (bf16 1.0 = 0x3f80)
    reinterpret_cast<uint4*>(sOnes)[threadIdx.x] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (wg == 4) {
    // The control warps are the highest-numbered warps of their SM sub-partitions: the warp scheduler favours the highest
    // warp id among eligible warps, and the MMA issuer must never queue behind four busy softmax warps.
    ptx::setmaxnreg_dec<32>();   // 128 x (96 - 32) registers released = 512 x (112 - 96) claimed below
    if (cw == 0) {
      // ---------------------------------------------------------- TMA producer
      // Ring order of the kv tiles: K_0, then for every j: K_{j+1} (if any), V_j.
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(q_full, 2 * FB_Q_BYTES);
        ptx::tma_load_3d(sQ, &tmQ, q_full, head * FB_D, m0, batch);
        ptx::tma_load_3d(sQ + FB_Q_BYTES, &tmQ, q_full, head * FB_D, m0 + FB_BM, batch);
        int slot = 0;
        uint32_t phase = 0;
        auto load = [&](const CUtensorMap* tm, int row0) {
          ptx::mbar_wait(&kv_empty[slot], phase ^ 1);
          ptx::mbar_expect_tx(&kv_full[slot], FB_TILE_BYTES);
          ptx::tma_load_3d(sKV + slot * FB_TILE_BYTES, tm, &kv_full[slot], head * FB_D, row0, batch);
          if (++slot == FB_SLOTS) { slot = 0; phase ^= 1; }
        };
        load(&tmK, 0);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) load(&tmK, (j + 1) * FB_BN);
          load(&tmV, j * FB_BN);
        }
      }
    } else {
      // ---------------------------------------------------------- tcgen05 issuers
      // One issuing thread per Q tile, so that no MMA group queues behind a wait that belongs to the other tile (a single
      // in-order issuer pays ~100 cycles of mbarrier latency per wait, eight waits per kv step, and was the bound of the
      // first version of this kernel). Every K / V ring slot is released by two commits (kv_empty counts 2), one per tile.
      constexpr uint32_t idesc_s = ptx::idesc_bf16(FB_BM, FB_BN, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = ptx::idesc_bf16(FB_BM, FB_D, 0, 1);   // P (TMEM) x V (MN-major)
      constexpr uint32_t idesc_l = ptx::idesc_bf16(FB_BM, 16, 0, 0);     // P (TMEM) x ones
      const uint32_t sKV_a = ptx::smem_u32(sKV);
      // ring position of a tile in the sequence K_0, K_1, V_0, K_2, V_1, ..., K_{n-1}, V_{n-2}, V_{n-1}
      auto idx_k = [&](int j) { return j == 0 ? 0 : 2 * j - 1; };
      auto idx_v = [&](int j) { return j < nkv - 1 ? 2 * j + 2 : 2 * nkv - 1; };
      auto wait_kv = [&](int idx) { ptx::mbar_wait(&kv_full[idx & (FB_SLOTS - 1)], (idx / FB_SLOTS) & 1); };
      if (cw <= 2 && ptx::elect_one()) {
        // control warp 1 + t issues everything of Q tile t, in the order its events occur:
        //   S_t(j+1) (K tile landed, S_t(j) read out by the softmax warps) -> PV_(t,0)(j) -> PV_(t,1)(j)   (V tile landed, P half published)
        const int t = cw - 1;
        const uint64_t a_desc = ptx::smem_desc_sw128(ptx::smem_u32(sQ) + t * FB_Q_BYTES, 16, 1024);
        const uint64_t ones_desc = ptx::smem_desc_sw128(ptx::smem_u32(sOnes), 16, 1024);
        const uint32_t tS = tmem_base + col_s(0) + 128 * t;
        const uint32_t tO = tmem_base + col_o(0) + 64 * t;
        const uint32_t tL = tmem_base + col_l(0) + 16 * t;
        auto do_s = [&](int j) {
          const int idx = idx_k(j);
          wait_kv(idx);
          if (j > 0) ptx::mbar_wait(&s_free[t], (j - 1) & 1);
          ptx::tc_fence_after();
          const uint64_t b = ptx::smem_desc_sw128(sKV_a + (idx & (FB_SLOTS - 1)) * FB_TILE_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < FB_D / 16; ++k)
            ptx::umma_ss(tS, a_desc + 2 * k, b + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          ptx::umma_commit(&s_full[t]);
          ptx::umma_commit(&kv_empty[idx & (FB_SLOTS - 1)]);
        };
        ptx::mbar_wait(q_full, 0);
        do_s(0);
        if (nkv > 1) do_s(1);
        // At the end of softmax tile j the warps first release S_t(j+1) and then publish P_t(j): S_t(j+2) goes first.
#pragma unroll 1
        for (int j = 0; j < nkv; ++j) {
          if (j + 2 < nkv) do_s(j + 2);
          const int idx = idx_v(j);
          wait_kv(idx);
          const uint64_t b0 = ptx::smem_desc_sw128(sKV_a + (idx & (FB_SLOTS - 1)) * FB_TILE_BYTES, 1024, 1024);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            ptx::mbar_wait(&p_ready[t * 2 + hh], j & 1);
            ptx::tc_fence_after();
            const uint32_t tp = tmem_base + (hh ? col_p(0, 1) + 32 * t : col_p(0, 0) + 128 * t);
#pragma unroll
            for (int kk = 0; kk < FB_KSTEPS; ++kk) {
              const uint32_t acc = (j == 0 && hh == 0 && kk == 0) ? 0u : 1u;
              ptx::umma_ts(tO, tp + kk * 8, b0 + (hh * FB_KSTEPS + kk) * 128, idesc_o, acc);
              ptx::umma_ts(tL, tp + kk * 8, ones_desc, idesc_l, acc);
            }
            ptx::umma_commit(&pv_done[t * 2 + hh]);
          }
          ptx::umma_commit(&kv_empty[idx & (FB_SLOTS - 1)]);
        }
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups
    ptx::setmaxnreg_inc<112>();
    const int t = wg >> 1;
    const int hh = wg & 1;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // row inside the Q tile
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = ptx::opaque(tmem_base + lane_addr + col_s(t) + hh * FB_HALF);
    const uint32_t tP = ptx::opaque(tmem_base + lane_addr + col_p(t, hh));
    const uint32_t a_s_full = ptx::opaque(ptx::smem_u32(&s_full[t]));
    const uint32_t a_s_free = ptx::opaque(ptx::smem_u32(&s_free[t]));
    const uint32_t a_p_ready = ptx::opaque(ptx::smem_u32(&p_ready[t * 2 + hh]));
    const uint32_t a_pv_done = ptx::opaque(ptx::smem_u32(&pv_done[t * 2 + hh]));
    const float sc = prm.scale_log2;
    const uint64_t c2 = f2_pack(sc, sc);
    const uint64_t negm2 = f2_pack(-m_off, -m_off);
    const uint64_t a2 = f2_pack(12582912.0f - m_off, 12582912.0f - m_off);
    const int tail = prm.Skv - (nkv - 1) * FB_BN - hh * FB_HALF;   // valid columns of this half in the last kv tile

    uint32_t s0[16], s1[16], s2[16];
    uint32_t pk[24];

    auto arrive_warp = [&](uint32_t bar) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_a(bar);
    };

    ptx::mbar_wait_a(a_s_full, 0);
    ptx::tc_fence_after();
    ptx::tmem_ld_32x16(tS, s0);
    ptx::tmem_ld_32x16(tS + 16, s1);
    ptx::tmem_ld_32x16(tS + 32, s2);
    ptx::tmem_ld_wait();
    arrive_warp(a_s_free);

    // One kv tile = three 16-column chunks. P stays in registers until the whole half row is done and goes back to TMEM in one
    // burst: p_ready is only published then anyway, and the PV product of tile j-1, which must have released the (single) P
    // buffer of this half, gets the whole tile to complete. Chunks 0 and 1 of the next S tile stream in under chunk 2; their
    // barrier is probed (test_wait: never suspends) one chunk early so that the probe latency is off the critical path.
    for (int j = 0; j < nkv; ++j) {
      const bool last = (j == nkv - 1);
      if (!last) exp_chunk<NP, false>(s0, &pk[0], c2, negm2, a2, 0, 0);
      else exp_chunk<NP, true>(s0, &pk[0], c2, negm2, a2, 0, tail);
      const bool s_ok = !last ? ptx::mbar_test_wait_a(a_s_full, (j + 1) & 1) : true;
      if (!last) exp_chunk<NP, false>(s1, &pk[8], c2, negm2, a2, 0, 0);
      else exp_chunk<NP, true>(s1, &pk[8], c2, negm2, a2, 16, tail);
      bool pv_ok = true;
      if (!last) {
        if (!s_ok) ptx::mbar_wait_a(a_s_full, (j + 1) & 1);
        ptx::tc_fence_after();
        ptx::tmem_ld_32x16(tS, s0);
        ptx::tmem_ld_32x16(tS + 16, s1);
        if (j > 0) pv_ok = ptx::mbar_test_wait_a(a_pv_done, (j - 1) & 1);
        exp_chunk<NP, false>(s2, &pk[16], c2, negm2, a2, 0, 0);
        ptx::tmem_ld_32x16(tS + 32, s2);
      } else {
        if (j > 0) pv_ok = ptx::mbar_test_wait_a(a_pv_done, (j - 1) & 1);
        exp_chunk<NP, true>(s2, &pk[16], c2, negm2, a2, 32, tail);
      }
      if (!pv_ok) ptx::mbar_wait_a(a_pv_done, (j - 1) & 1);     // P_(t,hh)(j-1) has been consumed
      ptx::tc_fence_after();
      ptx::tmem_st_32x16(tP, *reinterpret_cast<uint32_t (*)[16]>(&pk[0]));
      ptx::tmem_st_32x8(tP + 16, *reinterpret_cast<uint32_t (*)[8]>(&pk[16]));
      if (!last) {                                            // S_t(j+2) is wanted sooner than the PV product: release S first
        ptx::tmem_ld_wait();
        arrive_warp(a_s_free);
      }
      ptx::tmem_st_wait();
      arrive_warp(a_p_ready);
    }

    // ---------------------------------------------------------- epilogue: O / l -> bf16 global (32 of the 64 columns)
    ptx::mbar_wait(&pv_done[t * 2 + 0], (nkv - 1) & 1);
    ptx::mbar_wait(&pv_done[t * 2 + 1], (nkv - 1) & 1);
    ptx::tc_fence_after();
    uint32_t lv[16];
    ptx::tmem_ld_32x16(tmem_base + lane_addr + col_l(t), lv);
    uint32_t o0[16], o1[16];
    ptx::tmem_ld_32x16(tmem_base + lane_addr + col_o(t) + hh * 32, o0);
    ptx::tmem_ld_32x16(tmem_base + lane_addr + col_o(t) + hh * 32 + 16, o1);
    ptx::tmem_ld_wait();
    const float l = __uint_as_float(lv[0]);
    const float inv_l = 1.0f / l;
    const int row = m0 + t * FB_BM + r;
    if (row < prm.Sq) {
      if (prm.lse != nullptr && hh == 0)
        prm.lse[(static_cast<long long>(batch) * prm.H + head) * prm.Sq + row] = m_off + log2f(l);
      __nv_bfloat16* orow = prm.out + static_cast<long long>(batch) * prm.out_batch_stride +
                            static_cast<long long>(row) * prm.out_row_stride + head * FB_D + hh * 32;
      uint4 v;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        v.x = pack_bf16x2(__uint_as_float(o0[8 * c + 0]) * inv_l, __uint_as_float(o0[8 * c + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(o0[8 * c + 2]) * inv_l, __uint_as_float(o0[8 * c + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(o0[8 * c + 4]) * inv_l, __uint_as_float(o0[8 * c + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(o0[8 * c + 6]) * inv_l, __uint_as_float(o0[8 * c + 7]) * inv_l);
        reinterpret_cast<uint4*>(orow)[c] = v;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        v.x = pack_bf16x2(__uint_as_float(o1[8 * c + 0]) * inv_l, __uint_as_float(o1[8 * c + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(o1[8 * c + 2]) * inv_l, __uint_as_float(o1[8 * c + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(o1[8 * c + 4]) * inv_l, __uint_as_float(o1[8 * c + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(o1[8 * c + 6]) * inv_l, __uint_as_float(o1[8 * c + 7]) * inv_l);
        reinterpret_cast<uint4*>(orow)[2 + c] = v;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (cw == 1) ptx::tmem_dealloc(tmem_base, FB_TMEM_COLS);
}

// max |row|^2 per (batch, head) of q and of k: bounds[(b*H + h)*2 + {0: q, 1: k}] (atomicMax on the bits of a
// non-negative float; the buffer is zeroed by the caller). 8 lanes share one 64-element head row (16 bytes each).
constexpr int NB_ROWS = 512;   // rows per block

struct NbParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  long long q_row_stride, q_batch_stride, k_row_stride, k_batch_stride;
  int Sq, Skv, H, q_blocks;
  float* bounds;
};

__global__ void __launch_bounds__(256) attn_qk_bound_kernel(NbParams p) {
  const int head = blockIdx.y, batch = blockIdx.z;
  const bool is_k = static_cast<int>(blockIdx.x) >= p.q_blocks;
  const int blk = is_k ? blockIdx.x - p.q_blocks : blockIdx.x;
  const int S = is_k ? p.Skv : p.Sq;
  const __nv_bfloat16* base = (is_k ? p.k + batch * p.k_batch_stride : p.q + batch * p.q_batch_stride) + head * 64 + (threadIdx.x & 7) * 8;
  const long long rs = is_k ? p.k_row_stride : p.q_row_stride;
  const int r0 = blk * NB_ROWS + (threadIdx.x >> 3);
  float mx = 0.f;
#pragma unroll 4
  for (int i = 0; i < NB_ROWS / 32; ++i) {
    const int row = r0 + i * 32;
    float ss = 0.f;
    if (row < S) {
      const uint4 v = *reinterpret_cast<const uint4*>(base + row * rs);
      const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
      ss = a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    mx = fmaxf(mx, ss);
  }
  mx = warp_max(mx);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    // NaN / Inf rows must not pass as "small": map anything non-finite to +Inf (bits 0x7f800000 compare above all finite)
    if (!(mx <= 3.0e38f)) mx = INFINITY;
    atomicMax(reinterpret_cast<unsigned int*>(p.bounds) + 2 * (batch * p.H + head) + (is_k ? 1 : 0), __float_as_uint(mx));
  }
}

template <int NP>
int launch_fb(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const FbParams& prm, dim3 grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VGPA_CUDA(cudaFuncSetAttribute(attn_fwd_d64_bounded_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_BYTES));
    attr_set = true;
  }
  attn_fwd_d64_bounded_kernel<NP><<<grid, FB_THREADS, FB_SMEM_BYTES, stream>>>(tq, tk, tv, prm);
  VGPA_LAUNCH_CHECK("attn_fwd_d64_bounded_kernel");
  return 0;
}

}  // namespace

size_t attention_d64_workspace_bytes(int B, int H) { return static_cast<size_t>(B) * H * 2 * sizeof(float); }

// Pre-pass + bounded-softmax kernel. `bounds` (attention_d64_workspace_bytes) is left holding max|q|^2, max|k|^2 per
// (batch, head) so that the exact kernel, launched next by the caller, can skip the heads served here.
int launch_attention_d64_bounded(const vgpa_attention_args* a, float* bounds, float scale_log2, int npoly8, cudaStream_t stream) {
  const int cols = a->H * 64;
  VGPA_CUDA(cudaMemsetAsync(bounds, 0, attention_d64_workspace_bytes(a->B, a->H), stream));
  NbParams nb;
  nb.q = static_cast<const __nv_bfloat16*>(a->q);
  nb.k = static_cast<const __nv_bfloat16*>(a->k);
  nb.q_row_stride = a->q_row_stride; nb.q_batch_stride = a->q_batch_stride;
  nb.k_row_stride = a->k_row_stride; nb.k_batch_stride = a->k_batch_stride;
  nb.Sq = a->Sq; nb.Skv = a->Skv; nb.H = a->H;
  nb.q_blocks = (a->Sq + NB_ROWS - 1) / NB_ROWS;
  nb.bounds = bounds;
  dim3 ngrid(nb.q_blocks + (a->Skv + NB_ROWS - 1) / NB_ROWS, a->H, a->B);
  attn_qk_bound_kernel<<<ngrid, 256, 0, stream>>>(nb);
  VGPA_LAUNCH_CHECK("attn_qk_bound_kernel");

  CUtensorMap tq, tk, tv;
  {
    const uint32_t box[3] = {64, FB_BM, 1};
    const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)a->Sq, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->q_row_stride * 2, (uint64_t)a->q_batch_stride * 2};
    if (int rc = make_tmap_bf16(&tq, a->q, 3, dims, str, box)) return rc;
  }
  const uint32_t kvbox[3] = {64, FB_BN, 1};
  {
    const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)a->Skv, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->k_row_stride * 2, (uint64_t)a->k_batch_stride * 2};
    if (int rc = make_tmap_bf16(&tk, a->k, 3, dims, str, kvbox)) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)a->Skv, (uint64_t)a->B};
    const uint64_t str[2] = {(uint64_t)a->v_row_stride * 2, (uint64_t)a->v_batch_stride * 2};
    if (int rc = make_tmap_bf16(&tv, a->v, 3, dims, str, kvbox)) return rc;
  }
  FbParams prm;
  prm.out = static_cast<__nv_bfloat16*>(a->out);
  prm.out_row_stride = a->out_row_stride;
  prm.out_batch_stride = a->out_batch_stride;
  prm.Sq = a->Sq;
  prm.Skv = a->Skv;
  prm.scale_log2 = scale_log2;
  prm.lse = a->lse;
  prm.H = a->H;
  prm.bounds = bounds;
  dim3 grid((a->Sq + 2 * FB_BM - 1) / (2 * FB_BM), a->H, a->B);
  switch (npoly8) {
    case 0: return launch_fb<0>(tq, tk, tv, prm, grid, stream);
    case 2: return launch_fb<2>(tq, tk, tv, prm, grid, stream);
    case 3: return launch_fb<3>(tq, tk, tv, prm, grid, stream);
    case 5: return launch_fb<5>(tq, tk, tv, prm, grid, stream);
    default: return launch_fb<4>(tq, tk, tv, prm, grid, stream);
  }
}

}  // namespace vgpa
