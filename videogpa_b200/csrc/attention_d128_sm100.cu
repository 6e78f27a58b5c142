// K1b: full attention with head_dim 128 (Wan2.2 self- and cross-attention), same C entry point as the
// head_dim-64 kernel (vgpa_attention_bf16 dispatches on head_dim).
//
// Replaces flash_attention(q, k, v) inside Wan2.2's WanSelfAttention / WanCrossAttention (SURVEY.md App. A.7;
// reference call sites generate/Wan2.2-TI2V-5B.py:120-129, train/Wan2.2-TI2V-5B/03_train.py:228-233).
//
// With 128-wide heads there are 256 MMA FLOPs per softmax element, so tensor pipe and softmax are roughly
// balanced and the classic two-tile ping-pong is the right shape: one CTA = 256 query rows (2 x 128) of one
// (batch, head); warp 0 TMA producer, warp 1 tcgen05 issuer, warpgroups 1/2 = softmax of Q tile 0/1 (one
// thread per row). TMEM (512 columns, all used): S_t [t*128, +128), O_t [256 + t*128, +128). P_t (bf16) is
// written over the first 64 columns of S_t once the row is in registers and feeds a TS-form MMA; the issuer
// orders S_t(j+1) after PV_t(j), so no extra hand-shake is needed for the aliasing. While warpgroup t waits for
// PV_t(j) + S_t(j+1), the other warpgroup owns the MUFU/FMA pipes.
#include "sm100.cuh"
#include "attn_common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {
using namespace attn;

constexpr int A2_THREADS = 384;
constexpr int A2_BM = 128;
constexpr int A2_BN = 128;
constexpr int A2_D = 128;
constexpr int A2_SLOTS = 4;                                    // ring of 32 KB tiles: K0 V0 K1 V1 ...
constexpr uint32_t A2_HALF_BYTES = 128 * 64 * 2;               // one [128 x 64] swizzled block
constexpr uint32_t A2_TILE_BYTES = 2 * A2_HALF_BYTES;          // [128 x 128] as two column halves
constexpr uint32_t A2_SMEM_BYTES = 2 * A2_TILE_BYTES + A2_SLOTS * A2_TILE_BYTES + 1024 + 256;
constexpr uint32_t A2_TMEM_COLS = 512;
constexpr uint32_t A2_COL_O = 256;
constexpr float A2_RESCALE_THRESHOLD = 8.0f;
constexpr int A2_NPOLY = 16;

struct Attn2Params {
  __nv_bfloat16* out;
  long long out_row_stride;
  long long out_batch_stride;
  int Sq, Skv;
  float scale_log2;
};

__global__ void __launch_bounds__(A2_THREADS, 1)
attn_fwd_d128_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, Attn2Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + 2 * A2_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + A2_SLOTS * A2_TILE_BYTES);
  uint64_t* q_full = bars;                      // 1
  uint64_t* kv_full = bars + 1;                 // A2_SLOTS
  uint64_t* kv_empty = kv_full + A2_SLOTS;      // A2_SLOTS
  uint64_t* s_full = kv_empty + A2_SLOTS;       // [2]
  uint64_t* p_ready = s_full + 2;               // [2]
  uint64_t* pv_done = p_ready + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int m0 = blockIdx.x * (2 * A2_BM);
  const int nkv = (prm.Skv + A2_BN - 1) / A2_BN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < A2_SLOTS; ++i) {
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_ready[i], 128);
      ptx::mbar_init(&pv_done[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, A2_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (wg == 0) {
    ptx::setmaxnreg_dec<56>();   // 128*56 + 256*224 = 384*168: exactly the CTA's launch-time register allocation
    if (warp == 0) {
      // ---------------------------------------------------------- TMA producer (ring order K0 V0 K1 V1 ...)
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(q_full, 2 * A2_TILE_BYTES);
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf)
            ptx::tma_load_3d(sQ + t * A2_TILE_BYTES + hlf * A2_HALF_BYTES, &tmQ, q_full, head * A2_D + hlf * 64, m0 + t * A2_BM, batch);
        int slot = 0;
        uint32_t phase = 0;
        for (int j = 0; j < nkv; ++j) {
#pragma unroll
          for (int kv = 0; kv < 2; ++kv) {
            ptx::mbar_wait(&kv_empty[slot], phase ^ 1);
            ptx::mbar_expect_tx(&kv_full[slot], A2_TILE_BYTES);
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf)
              ptx::tma_load_3d(sKV + slot * A2_TILE_BYTES + hlf * A2_HALF_BYTES, kv == 0 ? &tmK : &tmV, &kv_full[slot],
                               head * A2_D + hlf * 64, j * A2_BN, batch);
            if (++slot == A2_SLOTS) { slot = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- tcgen05 issuer
      constexpr uint32_t idesc_s = ptx::idesc_bf16(A2_BM, A2_BN, 0, 0);   // Q (K-major) x K (K-major), K = head_dim
      constexpr uint32_t idesc_o = ptx::idesc_bf16(A2_BM, 64, 0, 1);      // P (TMEM) x V half (MN-major), N = 64
      const uint32_t sQ_a = ptx::smem_u32(sQ);
      const uint32_t sKV_a = ptx::smem_u32(sKV);
      if (ptx::elect_one()) {
        auto slot_of = [](int idx) { return idx % A2_SLOTS; };
        auto phase_of = [](int idx) { return static_cast<uint32_t>((idx / A2_SLOTS) & 1); };
        auto do_s = [&](int t, int kslot) {
#pragma unroll
          for (int k = 0; k < A2_D / 16; ++k) {
            const uint64_t a = ptx::smem_desc_sw128(sQ_a + t * A2_TILE_BYTES + (k >> 2) * A2_HALF_BYTES + (k & 3) * 32, 16, 1024);
            const uint64_t b = ptx::smem_desc_sw128(sKV_a + kslot * A2_TILE_BYTES + (k >> 2) * A2_HALF_BYTES + (k & 3) * 32, 16, 1024);
            ptx::umma_ss(tmem_base + t * A2_BN, a, b, idesc_s, k != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&s_full[t]);
        };
        auto do_pv = [&](int t, int vslot, bool first) {
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {                       // the two 64-column halves of O_t
#pragma unroll
            for (int kk = 0; kk < A2_BN / 16; ++kk) {
              const uint64_t b = ptx::smem_desc_sw128(sKV_a + vslot * A2_TILE_BYTES + hlf * A2_HALF_BYTES + kk * 2048, 1024, 1024);
              ptx::umma_ts(tmem_base + A2_COL_O + t * A2_D + hlf * 64, tmem_base + t * A2_BN + kk * 8, b, idesc_o,
                           (first && kk == 0) ? 0u : 1u);
            }
          }
          ptx::umma_commit(&pv_done[t]);
        };
        ptx::mbar_wait(q_full, 0);
        ptx::mbar_wait(&kv_full[0], 0);
        ptx::tc_fence_after();
        do_s(0, 0);
        do_s(1, 0);
        ptx::umma_commit(&kv_empty[0]);
        for (int j = 0; j < nkv; ++j) {
          const int vi = 2 * j + 1, ki = 2 * j + 2;
          const bool more = j + 1 < nkv;
          ptx::mbar_wait(&kv_full[slot_of(vi)], phase_of(vi));
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            ptx::mbar_wait(&p_ready[t], j & 1);
            ptx::tc_fence_after();
            do_pv(t, slot_of(vi), j == 0);
            if (t == 1) ptx::umma_commit(&kv_empty[slot_of(vi)]);
            if (more) {
              if (t == 0) {
                ptx::mbar_wait(&kv_full[slot_of(ki)], phase_of(ki));
                ptx::tc_fence_after();
              }
              do_s(t, slot_of(ki));                                  // ordered after PV_t(j): P_t aliases S_t
              if (t == 1) ptx::umma_commit(&kv_empty[slot_of(ki)]);
            }
          }
        }
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups
    ptx::setmaxnreg_inc<224>();
    const int t = wg - 1;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + t * A2_BN;
    const uint32_t tO = tmem_base + lane_addr + A2_COL_O + t * A2_D;
    const float sc = prm.scale_log2;
    const uint64_t sc2 = f2_pack(sc, sc);
    const int tail = prm.Skv - (nkv - 1) * A2_BN;
    float m_used = -INFINITY;
    uint64_t l2a = f2_pack(0.f, 0.f), l2b = l2a;

    for (int j = 0; j < nkv; ++j) {
      ptx::mbar_wait(&s_full[t], j & 1);
      ptx::tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld_32x32(tS + c * 32, *reinterpret_cast<uint32_t (*)[32]>(&s[c * 32]));
      ptx::tmem_ld_wait();
      if (j == nkv - 1 && tail < A2_BN) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= tail) s[i] = 0xff800000u;
      }
      float mx0 = max3(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
      float mx1 = max3(__uint_as_float(s[3]), __uint_as_float(s[4]), __uint_as_float(s[5]));
      float mx2 = max3(__uint_as_float(s[6]), __uint_as_float(s[7]), __uint_as_float(s[8]));
      float mx3 = max3(__uint_as_float(s[9]), __uint_as_float(s[10]), __uint_as_float(s[11]));
#pragma unroll
      for (int i = 12; i < 124; i += 8) {
        mx0 = max3(mx0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
        mx1 = max3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mx2 = max3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mx3 = max3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
      }
      mx0 = max3(mx0, __uint_as_float(s[124]), __uint_as_float(s[125]));
      mx1 = max3(mx1, __uint_as_float(s[126]), __uint_as_float(s[127]));
      const float m_cur = fmaxf(max3(mx0, mx1, mx2), mx3) * sc;
      const bool need = m_cur > m_used + A2_RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_used, m_cur);
        const float factor = ptx::ex2_approx(m_used - m_new);
        m_used = m_new;
        const uint64_t f2 = f2_pack(factor, factor);
        const uint64_t z2 = f2_pack(0.f, 0.f);
        l2a = f2_fma(l2a, f2, z2);
        l2b = f2_fma(l2b, f2, z2);
        if (j > 0) {
          // s_full(j) was committed after PV_t(j-1) on the issuing thread, so O_t is quiescent here
#pragma unroll
          for (int c = 0; c < A2_D / 16; ++c) {
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
      const uint64_t negm2 = f2_pack(-m_used, -m_used);
      // P_t(j) goes over the first 64 columns of S_t (every thread only touches its own TMEM lane; the whole
      // row is already in registers), 32 packed columns at a time to bound register pressure
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t pk[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int i = hh * 32 + q;
          const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc2, negm2);
          float p0, p1;
          if (pair_uses_poly<A2_NPOLY>(i)) {
            ex2_poly2(x2, p0, p1);
          } else {
            float x0, x1;
            f2_unpack(x2, x0, x1);
            p0 = ptx::ex2_approx(x0);
            p1 = ptx::ex2_approx(x1);
          }
          if (q & 1) l2b = f2_add(l2b, f2_pack(p0, p1)); else l2a = f2_add(l2a, f2_pack(p0, p1));
          pk[q] = pack_bf16x2(p0, p1);
        }
        ptx::tmem_st_32x32(tS + hh * 32, pk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_ready[t]);
    }

    // ---------------------------------------------------------- epilogue: O / l -> bf16 global
    ptx::mbar_wait(&pv_done[t], (nkv - 1) & 1);
    ptx::tc_fence_after();
    float la, lb, lc, ld;
    f2_unpack(l2a, la, lb);
    f2_unpack(l2b, lc, ld);
    const float inv_l = 1.0f / ((la + lb) + (lc + ld));
    const int row = m0 + t * A2_BM + r;
    __nv_bfloat16* orow = prm.out + static_cast<long long>(batch) * prm.out_batch_stride +
                          static_cast<long long>(row < prm.Sq ? row : 0) * prm.out_row_stride + head * A2_D;
#pragma unroll
    for (int c = 0; c < A2_D / 16; ++c) {
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
  if (warp == 1) ptx::tmem_dealloc(tmem_base, A2_TMEM_COLS);
}

}  // namespace

// called by vgpa_attention_bf16 (attention_sm100.cu) for head_dim 128; arguments are already validated
int launch_attention_d128(const vgpa_attention_args* a, cudaStream_t stream) {
  const int cols = a->H * 128;
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
  static bool attr_set = false;
  if (!attr_set) {
    VGPA_CUDA(cudaFuncSetAttribute(attn_fwd_d128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM_BYTES));
    attr_set = true;
  }
  Attn2Params prm;
  prm.out = static_cast<__nv_bfloat16*>(a->out);
  prm.out_row_stride = a->out_row_stride;
  prm.out_batch_stride = a->out_batch_stride;
  prm.Sq = a->Sq;
  prm.Skv = a->Skv;
  const float scale = a->scale > 0.f ? a->scale : 0.08838834764831845f;   // 1/sqrt(128)
  prm.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((a->Sq + 2 * A2_BM - 1) / (2 * A2_BM), a->H, a->B);
  attn_fwd_d128_kernel<<<grid, A2_THREADS, A2_SMEM_BYTES, stream>>>(tq, tk, tv, prm);
  VGPA_LAUNCH_CHECK("attn_fwd_d128_kernel");
  return 0;
}

}  // namespace vgpa
