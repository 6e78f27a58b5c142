// K1: joint text+video full attention, head_dim 64, for the CogVideoX DiT block.
//
// Replaces F.scaled_dot_product_attention inside diffusers' CogVideoXAttnProcessor2_0
// (reached from generate/CogVideoX-5B.py:72-77 and
// train/CogVideoX-5B/03_train.py:134-151). q/k arrive already LayerNorm'ed and rotated by the
// fused QKV GEMM epilogue (gemm_sm100.cu), so this kernel is softmax(q k^T / sqrt(d)) v only.
//
// One CTA = 256 query rows of one (batch, head): two 128-row Q tiles that ping-pong on the tensor
// pipe. 384 threads:
//   warpgroup 0: warp 0 = TMA producer (Q once, then K_j / V_j tiles through a 6-slot ring),
//                warp 1 = tcgen05 issuer (S_t = Q_t K_j^T and O_t += P_t V_j, all in TMEM),
//                warps 2-3 idle (they only donate registers)
//   warpgroup 1: softmax of Q tile 0 (one thread = one query row, straight out of TMEM lanes)
//   warpgroup 2: softmax of Q tile 1
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384).
// P_t is written by the softmax threads as bf16 into 128B-swizzled smem and consumed as the A
// operand of the PV MMA. The row max used for the exponentials is allowed to go stale by up to
// 2^8 (O/l are rescaled only when a row max grows by more than that), which keeps the TMEM
// read-modify-write of O off the steady-state path; the final O/l is exact.
#include "sm100.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {

constexpr int AT_THREADS = 384;
constexpr int AT_BM = 128;        // query rows per tile
constexpr int AT_BN = 128;        // kv rows per tile
constexpr int AT_D = 64;          // head dim
constexpr int AT_KV_SLOTS = 6;    // ring of 16 KB tiles: K0 V0 K1 V1 ...
constexpr uint32_t AT_TILE_BYTES = AT_BN * AT_D * 2;        // 16384
constexpr uint32_t AT_P_BYTES = AT_BM * AT_BN * 2;          // 32768
constexpr uint32_t AT_SMEM_BYTES = 2 * AT_TILE_BYTES + AT_KV_SLOTS * AT_TILE_BYTES + 2 * AT_P_BYTES + 1024 + 256;
constexpr uint32_t AT_TMEM_COLS = 512;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;  // log2 units

struct AttnParams {
  __nv_bfloat16* out;
  long long out_row_stride;
  long long out_batch_stride;
  int Sq, Skv;
  float scale_log2;
};

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_d64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, AttnParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + 2 * AT_TILE_BYTES;
  uint8_t* sP = sKV + AT_KV_SLOTS * AT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_P_BYTES);
  uint64_t* q_full = bars;                        // 1
  uint64_t* kv_full = bars + 1;                   // AT_KV_SLOTS
  uint64_t* kv_empty = kv_full + AT_KV_SLOTS;     // AT_KV_SLOTS
  uint64_t* s_full = kv_empty + AT_KV_SLOTS;      // 2
  uint64_t* p_ready = s_full + 2;                 // 2
  uint64_t* o_final = p_ready + 2;                // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int m0 = blockIdx.x * (2 * AT_BM);
  const int nkv = (prm.Skv + AT_BN - 1) / AT_BN;

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
      ptx::mbar_init(&p_ready[i], 128);
      ptx::mbar_init(&o_final[i], 1);
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
    if (warp == 0) {
      // ---------------------------------------------------------- TMA producer
      if (lane == 0) {
        ptx::mbar_expect_tx(q_full, 2 * AT_TILE_BYTES);
        ptx::tma_load_3d(sQ, &tmQ, q_full, head * AT_D, m0, batch);
        ptx::tma_load_3d(sQ + AT_TILE_BYTES, &tmQ, q_full, head * AT_D, m0 + AT_BM, batch);
        int slot = 0;
        uint32_t phase = 0;
        for (int j = 0; j < nkv; ++j) {
#pragma unroll
          for (int kv = 0; kv < 2; ++kv) {
            ptx::mbar_wait(&kv_empty[slot], phase ^ 1);
            ptx::mbar_expect_tx(&kv_full[slot], AT_TILE_BYTES);
            ptx::tma_load_3d(sKV + slot * AT_TILE_BYTES, kv == 0 ? &tmK : &tmV, &kv_full[slot],
                             head * AT_D, j * AT_BN, batch);
            if (++slot == AT_KV_SLOTS) { slot = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- tcgen05 issuer
      constexpr uint32_t idesc_s = ptx::idesc_bf16(AT_BM, AT_BN, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = ptx::idesc_bf16(AT_BM, AT_D, 0, 1);   // P (K-major) x V (MN-major)
      const uint32_t sQ_a = ptx::smem_u32(sQ);
      const uint32_t sKV_a = ptx::smem_u32(sKV);
      const uint32_t sP_a = ptx::smem_u32(sP);
      int slot = 0;
      uint32_t phase = 0;
      auto issue_s = [&](int t, int kslot) {
        if (lane == 0) {
          const uint64_t a = ptx::smem_desc_sw128(sQ_a + t * AT_TILE_BYTES, 16, 1024);
          const uint64_t b = ptx::smem_desc_sw128(sKV_a + kslot * AT_TILE_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k)
            ptx::umma_ss(tmem_base + t * AT_BN, a + 2 * k, b + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          ptx::umma_commit(&s_full[t]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int vslot, bool accumulate) {
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < AT_BN / 16; ++k) {
            const uint64_t a = ptx::smem_desc_sw128(
                sP_a + t * AT_P_BYTES + (k >> 2) * (AT_BM * 128) + (k & 3) * 32, 16, 1024);
            const uint64_t b = ptx::smem_desc_sw128(sKV_a + vslot * AT_TILE_BYTES + k * 2048, 1024, 1024);
            ptx::umma_ss(tmem_base + 2 * AT_BN + t * AT_D, a, b, idesc_o, (accumulate || k != 0) ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto advance = [&]() { if (++slot == AT_KV_SLOTS) { slot = 0; phase ^= 1; } };

      ptx::mbar_wait(q_full, 0);
      // prologue: S_0(0), S_1(0)
      ptx::mbar_wait(&kv_full[slot], phase);
      ptx::tc_fence_after();
      issue_s(0, slot);
      issue_s(1, slot);
      if (lane == 0) ptx::umma_commit(&kv_empty[slot]);
      __syncwarp();
      advance();
      for (int j = 0; j < nkv; ++j) {
        const int vslot = slot;
        const uint32_t vphase = phase;
        advance();
        const int kslot = slot;          // K_{j+1}
        const uint32_t kphase = phase;
        const bool more = (j + 1 < nkv);
        if (more) advance();
        ptx::mbar_wait(&kv_full[vslot], vphase);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          ptx::mbar_wait(&p_ready[t], j & 1);
          ptx::tc_fence_after();
          issue_pv(t, vslot, j > 0);
          if (lane == 0) {
            if (t == 1) ptx::umma_commit(&kv_empty[vslot]);
            if (!more) ptx::umma_commit(&o_final[t]);
          }
          __syncwarp();
          if (more) {
            if (t == 0) {
              ptx::mbar_wait(&kv_full[kslot], kphase);
              ptx::tc_fence_after();
            }
            issue_s(t, kslot);
            if (t == 1) {
              if (lane == 0) ptx::umma_commit(&kv_empty[kslot]);
              __syncwarp();
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups
    ptx::setmaxnreg_inc<216>();
    const int t = wg - 1;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // row inside the Q tile
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + t * AT_BN;
    const uint32_t tO = tmem_base + lane_addr + 2 * AT_BN + t * AT_D;
    const uint32_t p_row = ptx::smem_u32(sP) + t * AT_P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
    const uint32_t xr = r & 7;
    const float sc = prm.scale_log2;
    const int tail = prm.Skv - (nkv - 1) * AT_BN;            // valid kv rows in the last tile
    float m_used = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < nkv; ++j) {
      ptx::mbar_wait(&s_full[t], j & 1);
      ptx::tc_fence_after();
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld_32x32(tS + c * 32, s[c]);
      ptx::tmem_ld_wait();
      if (j == nkv - 1 && tail < AT_BN) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= tail) s[c][i] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(s[0][i]));
        mx1 = fmaxf(mx1, __uint_as_float(s[1][i]));
        mx2 = fmaxf(mx2, __uint_as_float(s[2][i]));
        mx3 = fmaxf(mx3, __uint_as_float(s[3][i]));
      }
      const float m_cur = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
      const bool need = m_cur > m_used + AT_RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_used, m_cur);
        const float factor = ptx::ex2_approx(m_used - m_new);   // j == 0: exp2(-inf) = 0
        m_used = m_new;
        l *= factor;
        if (j > 0) {
          // s_full(j) was committed after PV(j-1) on the same issuing thread, so O is quiescent here
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
      const float neg_m = -m_used;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float p[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            p[i] = ptx::ex2_approx(fmaf(__uint_as_float(s[c][g * 8 + i]), sc, neg_m));
          l0 += p[0] + p[4];
          l1 += p[1] + p[5];
          l2 += p[2] + p[6];
          l3 += p[3] + p[7];
          const int chunk = c * 4 + g;                     // 16-byte chunk index along the row (0..15)
          const uint32_t addr = p_row + (chunk >> 3) * (AT_BM * 128) + (((chunk & 7) ^ xr) << 4);
          ptx::st_shared_v4(addr, pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]),
                            pack_bf16x2(p[4], p[5]), pack_bf16x2(p[6], p[7]));
        }
      }
      l += (l0 + l1) + (l2 + l3);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_ready[t]);
    }

    // ---------------------------------------------------------- epilogue: O / l -> bf16 global
    ptx::mbar_wait(&o_final[t], 0);
    ptx::tc_fence_after();
    const int row = m0 + t * AT_BM + r;
    const float inv_l = 1.0f / l;
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

}  // namespace
}  // namespace vgpa

extern "C" int vgpa_attention_bf16(const vgpa_attention_args* a, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr, "vgpa_attention_bf16: null args");
  VGPA_CHECK(a->head_dim == 64, "vgpa_attention_bf16: only head_dim 64 is built (got %d)", a->head_dim);
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
  const int cols = a->H * 64;
  VGPA_CHECK(a->q_row_stride >= cols && a->k_row_stride >= cols && a->v_row_stride >= cols && a->out_row_stride >= cols,
             "vgpa_attention_bf16: row strides must cover H*64 columns");
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
    VGPA_CUDA(cudaFuncSetAttribute(attn_fwd_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES));
    attr_set = true;
  }
  AttnParams prm;
  prm.out = static_cast<__nv_bfloat16*>(a->out);
  prm.out_row_stride = a->out_row_stride;
  prm.out_batch_stride = a->out_batch_stride;
  prm.Sq = a->Sq;
  prm.Skv = a->Skv;
  const float scale = a->scale > 0.f ? a->scale : 0.125f;
  prm.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((a->Sq + 2 * AT_BM - 1) / (2 * AT_BM), a->H, a->B);
  attn_fwd_d64_kernel<<<grid, AT_THREADS, AT_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(tq, tk, tv, prm);
  VGPA_LAUNCH_CHECK("attn_fwd_d64_kernel");
  return 0;
}
