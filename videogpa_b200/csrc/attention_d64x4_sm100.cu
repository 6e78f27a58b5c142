// K1 (head_dim 64), four-warpgroup softmax variant of attention_sm100.cu.
//
// Same math, same TMEM plan and same TS-form PV as attn_fwd_d64_kernel, but the softmax of one 128-row Q tile is split
// over TWO warpgroups by kv column half (a thread owns one query row x 64 kv columns), so every SM sub-partition runs
// four softmax warps instead of two. With head_dim 64 the kernel is bound by MUFU / FMA issue in the softmax warps, and
// with only two warps per scheduler their dependent-issue stalls were exposed (ncu: issue slots 54 % busy, XU 59 %,
// profiles/r01_attn_ncu_summary.md); four warps hide them.
//
// The two threads that share a row agree on the row maximum every kv tile through shared memory and a 64-thread named
// barrier (both warps live on the same sub-partition: TMEM lane quarter = warp % 4), so the lazy-rescale decision is exact
// and identical on both; each rescales its half of the O columns; the row sums are merged in the epilogue.
//
// 640 threads: warpgroup 0 = TMA producer + tcgen05 issuer, warpgroup 1 + 2t + hh = softmax of Q tile t, kv half hh.
#include "sm100.cuh"
#include "attn_common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {
using namespace attn;

constexpr int X4_THREADS = 640;
constexpr int X4_BM = 128;
constexpr int X4_BN = 128;
constexpr int X4_D = 64;
constexpr int X4_SLOTS = 6;
constexpr uint32_t X4_TILE_BYTES = X4_BN * X4_D * 2;   // 16384
constexpr uint32_t X4_XCH_BYTES = 2 * 2 * 2 * 128 * 4;  // [parity][tile][half][row] floats
constexpr uint32_t X4_SMEM_BYTES = 2 * X4_TILE_BYTES + X4_SLOTS * X4_TILE_BYTES + X4_XCH_BYTES + 1024 + 256;
constexpr uint32_t X4_TMEM_COLS = 512;
constexpr uint32_t X4_COL_P = 256;
constexpr uint32_t X4_COL_O = 384;
constexpr float X4_RESCALE_THRESHOLD = 8.0f;

struct X4Params {
  __nv_bfloat16* out;
  long long out_row_stride;
  long long out_batch_stride;
  int Sq, Skv;
  float scale_log2;
};

__device__ __forceinline__ void pair_barrier(int id) {
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

template <int NPOLY>
__global__ void __launch_bounds__(X4_THREADS, 1)
attn_fwd_d64x4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, X4Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + 2 * X4_TILE_BYTES;
  float* xch = reinterpret_cast<float*>(sKV + X4_SLOTS * X4_TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xch) + X4_XCH_BYTES);
  uint64_t* q_full = bars;                        // 1
  uint64_t* kv_full = bars + 1;                   // X4_SLOTS
  uint64_t* kv_empty = kv_full + X4_SLOTS;        // X4_SLOTS
  uint64_t* s_full = kv_empty + X4_SLOTS;         // [2]
  uint64_t* s_free = s_full + 2;                  // [2]    256 arrivals: both column halves hold S_t(j) in registers
  uint64_t* p_ready = s_free + 2;                 // [2][2] 128 arrivals
  uint64_t* pv_done = p_ready + 4;                // [2][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int m0 = blockIdx.x * (2 * X4_BM);
  const int nkv = (prm.Skv + X4_BN - 1) / X4_BN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < X4_SLOTS; ++i) {
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_free[i], 256);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&p_ready[i], 128);
      ptx::mbar_init(&pv_done[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, X4_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    ptx::setmaxnreg_dec<32>();      // 128*32 + 512*112 = 640*96: the CTA's launch-time register allocation
    // Ring order of the 16 KB tiles: K_0, then for every j: K_{j+1} (if any), V_j.
    if (warp == 0) {
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(q_full, 2 * X4_TILE_BYTES);
        ptx::tma_load_3d(sQ, &tmQ, q_full, head * X4_D, m0, batch);
        ptx::tma_load_3d(sQ + X4_TILE_BYTES, &tmQ, q_full, head * X4_D, m0 + X4_BM, batch);
        int slot = 0;
        uint32_t phase = 0;
        auto load = [&](const CUtensorMap* tm, int row0) {
          ptx::mbar_wait(&kv_empty[slot], phase ^ 1);
          ptx::mbar_expect_tx(&kv_full[slot], X4_TILE_BYTES);
          ptx::tma_load_3d(sKV + slot * X4_TILE_BYTES, tm, &kv_full[slot], head * X4_D, row0, batch);
          if (++slot == X4_SLOTS) { slot = 0; phase ^= 1; }
        };
        load(&tmK, 0);
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) load(&tmK, (j + 1) * X4_BN);
          load(&tmV, j * X4_BN);
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc_s = ptx::idesc_bf16(X4_BM, X4_BN, 0, 0);
      constexpr uint32_t idesc_o = ptx::idesc_bf16(X4_BM, X4_D, 0, 1);
      const uint32_t sQ_a = ptx::smem_u32(sQ);
      const uint32_t sKV_a = ptx::smem_u32(sKV);
      if (ptx::elect_one()) {
        auto idx_k = [&](int j) { return j == 0 ? 0 : 2 * j - 1; };
        auto idx_v = [&](int j) { return j < nkv - 1 ? 2 * j + 2 : 2 * nkv - 1; };
        auto wait_kv = [&](int idx) { ptx::mbar_wait(&kv_full[idx % X4_SLOTS], (idx / X4_SLOTS) & 1); };
        auto kv_release = [&](int idx) { ptx::umma_commit(&kv_empty[idx % X4_SLOTS]); };
        auto do_s = [&](int t, int idx) {
          const uint64_t a = ptx::smem_desc_sw128(sQ_a + t * X4_TILE_BYTES, 16, 1024);
          const uint64_t b = ptx::smem_desc_sw128(sKV_a + (idx % X4_SLOTS) * X4_TILE_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < X4_D / 16; ++k)
            ptx::umma_ss(tmem_base + t * X4_BN, a + 2 * k, b + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          ptx::umma_commit(&s_full[t]);
        };
        auto pv = [&](int t, int hh, int j) {
          ptx::mbar_wait(&p_ready[t * 2 + hh], j & 1);
          ptx::tc_fence_after();
          const int idx = idx_v(j);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t b = ptx::smem_desc_sw128(sKV_a + (idx % X4_SLOTS) * X4_TILE_BYTES + (hh * 4 + kk) * 2048, 1024, 1024);
            ptx::umma_ts(tmem_base + X4_COL_O + t * X4_D, tmem_base + X4_COL_P + t * 64 + hh * 32 + kk * 8, b, idesc_o,
                         (j == 0 && hh == 0 && kk == 0) ? 0u : 1u);
          }
          ptx::umma_commit(&pv_done[t * 2 + hh]);
        };
        auto sq = [&](int t, int j) {   // S_t(j+1)
          ptx::mbar_wait(&s_free[t], j & 1);
          ptx::tc_fence_after();
          do_s(t, idx_k(j + 1));
        };
        ptx::mbar_wait(q_full, 0);
        wait_kv(0);
        ptx::tc_fence_after();
        do_s(0, 0);
        do_s(1, 0);
        kv_release(0);
        // tile 0's PV(j, half 0) must be the very first accumulation into O_0; same for tile 1 (half 0 before half 1)
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) { wait_kv(idx_k(j + 1)); sq(0, j); }
          if (j > 0) { pv(1, 0, j - 1); pv(1, 1, j - 1); kv_release(idx_v(j - 1)); }
          if (j + 1 < nkv) { sq(1, j); kv_release(idx_k(j + 1)); }
          wait_kv(idx_v(j));
          pv(0, 0, j);
          pv(0, 1, j);
        }
        pv(1, 0, nkv - 1);
        pv(1, 1, nkv - 1);
        kv_release(idx_v(nkv - 1));
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups
    ptx::setmaxnreg_inc<112>();
    const int sw = warp - 4;
    const int t = sw >> 3;
    const int hh = (sw >> 2) & 1;
    const int quarter = sw & 3;                               // == warp & 3: the TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + t * X4_BN + hh * 64;
    const uint32_t tP = tmem_base + lane_addr + X4_COL_P + t * 64 + hh * 32;
    const uint32_t tO = tmem_base + lane_addr + X4_COL_O + t * X4_D + hh * 32;     // this thread's half of the O columns
    const int bar_id = 1 + t * 4 + quarter;
    const float sc = prm.scale_log2;
    const uint64_t sc2 = f2_pack(sc, sc);
    const int tail = prm.Skv - (nkv - 1) * X4_BN;
    float m_used = -INFINITY;
    uint64_t l2a = f2_pack(0.f, 0.f), l2b = l2a;

    for (int j = 0; j < nkv; ++j) {
      ptx::mbar_wait(&s_full[t], j & 1);
      ptx::tc_fence_after();
      uint32_t s[64];
      ptx::tmem_ld_32x32(tS, *reinterpret_cast<uint32_t (*)[32]>(&s[0]));
      ptx::tmem_ld_32x32(tS + 32, *reinterpret_cast<uint32_t (*)[32]>(&s[32]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&s_free[t]);
      if (j == nkv - 1 && tail < X4_BN) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (hh * 64 + i >= tail) s[i] = 0xff800000u;
      }
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
      const float hmax = fmaxf(max3(mx0, mx1, mx2), mx3);
      // exact row maximum: exchange with the thread that owns the other 64 columns of this row
      float* slot = xch + ((j & 1) * 4 + t * 2) * 128;
      slot[hh * 128 + r] = hmax;
      pair_barrier(bar_id);
      const float m_cur = fmaxf(hmax, slot[(hh ^ 1) * 128 + r]) * sc;
      const bool need = m_cur > m_used + X4_RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {                     // same rows => same decision in the partner warp
        const float m_new = fmaxf(m_used, m_cur);
        const float factor = ptx::ex2_approx(m_used - m_new);
        m_used = m_new;
        const uint64_t f2 = f2_pack(factor, factor);
        const uint64_t z2 = f2_pack(0.f, 0.f);
        l2a = f2_fma(l2a, f2, z2);
        l2b = f2_fma(l2b, f2, z2);
        if (j > 0) {
          // both PVs of tile j-1 add into every O column: wait for both, rescale my 32 columns, and make sure the
          // partner has rescaled its 32 before either half publishes P(j)
          ptx::mbar_wait(&pv_done[t * 2 + 0], (j - 1) & 1);
          ptx::mbar_wait(&pv_done[t * 2 + 1], (j - 1) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[16];
            ptx::tmem_ld_32x16(tO + c * 16, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            ptx::tmem_st_32x16(tO + c * 16, o);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          pair_barrier(bar_id);
        }
      }
      const uint64_t negm2 = f2_pack(-m_used, -m_used);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, negm2);
        float p0, p1;
        if (pair_uses_poly<NPOLY>(i * 2 + 1)) {                  // NPOLY of 64 pair slots; odd slots so that NPOLY = 16 -> 8 of the 32 pairs
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
      if (j > 0) {                                               // P_t(j-1, hh) has been consumed
        ptx::mbar_wait(&pv_done[t * 2 + hh], (j - 1) & 1);
        ptx::tc_fence_after();
      }
      ptx::tmem_st_32x32(tP, pk);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_ready[t * 2 + hh]);
    }

    // ---------------------------------------------------------- epilogue: merge the two row sums, O / l -> bf16 global
    float la, lb, lc, ld;
    f2_unpack(l2a, la, lb);
    f2_unpack(l2b, lc, ld);
    const float l_half = (la + lb) + (lc + ld);
    float* slot = xch + ((nkv & 1) * 4 + t * 2) * 128;
    slot[hh * 128 + r] = l_half;
    pair_barrier(bar_id);
    const float inv_l = 1.0f / (l_half + slot[(hh ^ 1) * 128 + r]);
    ptx::mbar_wait(&pv_done[t * 2 + 0], (nkv - 1) & 1);
    ptx::mbar_wait(&pv_done[t * 2 + 1], (nkv - 1) & 1);
    ptx::tc_fence_after();
    const int row = m0 + t * X4_BM + r;
    __nv_bfloat16* orow = prm.out + static_cast<long long>(batch) * prm.out_batch_stride +
                          static_cast<long long>(row < prm.Sq ? row : 0) * prm.out_row_stride + head * X4_D + hh * 32;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
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
  if (warp == 1) ptx::tmem_dealloc(tmem_base, X4_TMEM_COLS);
}

template <int NPOLY>
int launch_x4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const X4Params& prm, dim3 grid, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VGPA_CUDA(cudaFuncSetAttribute(attn_fwd_d64x4_kernel<NPOLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, X4_SMEM_BYTES));
    attr_set = true;
  }
  attn_fwd_d64x4_kernel<NPOLY><<<grid, X4_THREADS, X4_SMEM_BYTES, stream>>>(tq, tk, tv, prm);
  VGPA_LAUNCH_CHECK("attn_fwd_d64x4_kernel");
  return 0;
}

}  // namespace

// called by vgpa_attention_bf16 (attention_sm100.cu) for head_dim 64; arguments are validated and the maps built there
int launch_attention_d64x4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const vgpa_attention_args* a,
                           int npoly, cudaStream_t stream) {
  X4Params prm;
  prm.out = static_cast<__nv_bfloat16*>(a->out);
  prm.out_row_stride = a->out_row_stride;
  prm.out_batch_stride = a->out_batch_stride;
  prm.Sq = a->Sq;
  prm.Skv = a->Skv;
  const float scale = a->scale > 0.f ? a->scale : 0.125f;
  prm.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((a->Sq + 2 * X4_BM - 1) / (2 * X4_BM), a->H, a->B);
  switch (npoly) {
    case 0: return launch_x4<0>(tq, tk, tv, prm, grid, stream);
    case 8: return launch_x4<8>(tq, tk, tv, prm, grid, stream);
    case 24: return launch_x4<24>(tq, tk, tv, prm, grid, stream);
    case 32: return launch_x4<32>(tq, tk, tv, prm, grid, stream);
    default: return launch_x4<16>(tq, tk, tv, prm, grid, stream);
  }
}

}  // namespace vgpa
