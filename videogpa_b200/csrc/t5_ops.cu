// T5 (v1.1, umT5-style) encoder helpers — SURVEY.md §8 row f-4: the prompt encoder the reference calls as
// `pipeline.text_encoder(input_ids)[0]` (train/CogVideoX-5B/02_encode.py:69-84; 226 tokens, no attention mask).
//   * self-attention with the additive relative-position bias, no 1/sqrt(d) scaling (transformers T5Attention):
//     scores = bf16(q k^T); scores = bf16(scores + bias); softmax in fp32; P rounded to bf16; P v.
//     Sequences are short (226 tokens): 0.84 GFLOP per layer, so this is a shared-memory CUDA-core kernel, one CTA per
//     (head, sample, 32-query block) with the head's K and V resident in shared memory.
//   * the gated-GELU product hidden_gelu * hidden_linear of T5DenseGatedActDense.
// The projections run on vgpa_linear_bf16 and T5LayerNorm on vgpa_rmsnorm_rope_bf16.
#include "common.cuh"
#include "../../include/videogpa_b200.h"
#include <stdlib.h>

namespace vgpa {
namespace {

constexpr int T5_D = 64;            // d_kv
constexpr int T5_PAD = 66;          // padded row (bf16) so that lanes reading different keys hit different banks
constexpr int T5_QB = 32;           // queries per CTA
constexpr int T5_WARPS = 8;
constexpr int T5_SMAX = 512;
constexpr int T5_KPL = T5_SMAX / 32;   // keys per lane

struct T5AttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  const __nv_bfloat16* bias;   // [H, S, S]
  __nv_bfloat16* out;
  long long ld_qkv, ldo;
  int S, H;
};

__global__ void __launch_bounds__(T5_WARPS * 32)
t5_attention_kernel(T5AttnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_raw);             // [S][T5_PAD]
  __nv_bfloat16* Vs = Ks + static_cast<size_t>(p.S) * T5_PAD;                  // [S][T5_PAD]
  float* Ps = reinterpret_cast<float*>(Vs + static_cast<size_t>(p.S) * T5_PAD); // [T5_WARPS][S]
  const int h = blockIdx.x, b = blockIdx.y, q0 = blockIdx.z * T5_QB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = static_cast<long long>(b) * p.S;
  // stage K and V of this head: 8 threads per row (8 x 16 bytes = 64 bf16)
  for (int i = threadIdx.x; i < p.S * 8; i += T5_WARPS * 32) {
    const int r = i >> 3, c = (i & 7) * 8;
    const uint4 ku = *reinterpret_cast<const uint4*>(p.k + (row0 + r) * p.ld_qkv + h * T5_D + c);
    const uint4 vu = *reinterpret_cast<const uint4*>(p.v + (row0 + r) * p.ld_qkv + h * T5_D + c);
    uint32_t* kd = reinterpret_cast<uint32_t*>(Ks + r * T5_PAD + c);           // 4-byte aligned (T5_PAD even)
    uint32_t* vd = reinterpret_cast<uint32_t*>(Vs + r * T5_PAD + c);
    kd[0] = ku.x; kd[1] = ku.y; kd[2] = ku.z; kd[3] = ku.w;
    vd[0] = vu.x; vd[1] = vu.y; vd[2] = vu.z; vd[3] = vu.w;
  }
  __syncthreads();
  float* pw = Ps + static_cast<size_t>(warp) * p.S;
  for (int qi = warp; qi < T5_QB; qi += T5_WARPS) {
    const int qrow = q0 + qi;
    if (qrow >= p.S) break;
    // the query row, broadcast to every lane as 32 packed pairs
    uint32_t qreg[T5_D / 2];
    {
      const uint4* qp = reinterpret_cast<const uint4*>(p.q + (row0 + qrow) * p.ld_qkv + h * T5_D);
#pragma unroll
      for (int i = 0; i < T5_D / 8; ++i) {
        const uint4 u = __ldg(qp + i);
        qreg[4 * i] = u.x; qreg[4 * i + 1] = u.y; qreg[4 * i + 2] = u.z; qreg[4 * i + 3] = u.w;
      }
    }
    const __nv_bfloat16* brow = p.bias + (static_cast<long long>(h) * p.S + qrow) * p.S;
    float s[T5_KPL];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < T5_KPL; ++t) {
      const int j = t * 32 + lane;
      s[t] = -INFINITY;
      if (j < p.S) {
        const uint32_t* kr = reinterpret_cast<const uint32_t*>(Ks + j * T5_PAD);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < T5_D / 2; ++i) {
          const float2 kf = unpack_bf16x2(kr[i]), qf = unpack_bf16x2(qreg[i]);
          acc = fmaf(qf.x, kf.x, acc);
          acc = fmaf(qf.y, kf.y, acc);
        }
        s[t] = bf16_round(bf16_round(acc) + __bfloat162float(brow[j]));        // matmul output, then `scores += position_bias`
        mx = fmaxf(mx, s[t]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < T5_KPL; ++t) {
      const int j = t * 32 + lane;
      if (j < p.S) { s[t] = __expf(s[t] - mx); sum += s[t]; }
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < T5_KPL; ++t) {
      const int j = t * 32 + lane;
      if (j < p.S) pw[j] = bf16_round(s[t] * inv);                             // softmax(...).type_as(scores)
    }
    __syncwarp();
    // out[qrow, 2*lane .. 2*lane+1] = sum_j P[j] * V[j, :]
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < p.S; ++j) {
      const float pj = pw[j];
      const float2 vf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(Vs + j * T5_PAD + 2 * lane));
      o0 = fmaf(pj, vf.x, o0);
      o1 = fmaf(pj, vf.y, o1);
    }
    *reinterpret_cast<uint32_t*>(p.out + (row0 + qrow) * p.ldo + h * T5_D + 2 * lane) = pack_bf16x2(o0, o1);
    __syncwarp();
  }
}

// ---- the same attention on the tensor cores, for S <= 256 (the prompt encoder's 226 tokens).
// The CUDA-core kernel above spends 107 us per layer at S = 226 (35 % of the encoder: every FMA pays a bf16 unpack and a
// shared-memory read). Here one warp owns 16 query rows and runs both contractions as legacy warp-level MMAs
// (mma.sync.m16n8k16 bf16 -> fp32; 0.84 GFLOP per layer does not justify a tcgen05 pipeline): scores S = Q K^T for all keys stay in
// the accumulator registers (S / 8 tiles of 16x8), bias / rounding / softmax act on that layout (a row lives in one quad: two
// shuffles per reduction), and the rounded P tiles ARE the A fragments of the P V product (accumulator tiles 2i, 2i+1 = the
// 16 keys of k-step i). Rounding points are those of transformers' T5Attention, as in the kernel above.
constexpr int T5M_LD = 72;          // padded K / V row in shared memory (bf16): 144-byte stride, fragment loads hit 32 distinct banks
constexpr int T5M_NT = 32;          // key tiles of 8 held in registers: S <= 256
constexpr int T5M_WARPS = 4;        // 64 query rows per CTA

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(T5M_WARPS * 32)
t5_attention_mma_kernel(T5AttnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Spad = (p.S + 15) & ~15;
  const int ntiles = Spad >> 3;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_raw);             // [Spad][T5M_LD]
  __nv_bfloat16* Vs = Ks + static_cast<size_t>(Spad) * T5M_LD;                 // [Spad][T5M_LD]
  const int h = blockIdx.x, b = blockIdx.y, q0 = blockIdx.z * (T5M_WARPS * 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;                                       // fragment row group / column pair
  const long long row0 = static_cast<long long>(b) * p.S;
  // stage K and V of this head (rows >= S are zero): 8 threads per row, 16 bytes each
#pragma unroll 4
  for (int i = threadIdx.x; i < Spad * 8; i += T5M_WARPS * 32) {
    const int r = i >> 3, cc = (i & 7) * 8;
    uint4 ku = make_uint4(0, 0, 0, 0), vu = ku;
    if (r < p.S) {
      ku = *reinterpret_cast<const uint4*>(p.k + (row0 + r) * p.ld_qkv + h * T5_D + cc);
      vu = *reinterpret_cast<const uint4*>(p.v + (row0 + r) * p.ld_qkv + h * T5_D + cc);
    }
    *reinterpret_cast<uint4*>(Ks + r * T5M_LD + cc) = ku;                      // 144-byte rows: 16-byte aligned
    *reinterpret_cast<uint4*>(Vs + r * T5M_LD + cc) = vu;
  }
  __syncthreads();
  const int qa = q0 + warp * 16 + g, qb = qa + 8;                              // the two query rows of this thread
  if (q0 + warp * 16 >= p.S) return;
  const int qa_c = min(qa, p.S - 1), qb_c = min(qb, p.S - 1);                  // clamped for loads; stores are guarded
  // Q fragments: 4 k-steps of 16 dims
  uint32_t aq[4][4];
  {
    const __nv_bfloat16* qra = p.q + (row0 + qa_c) * p.ld_qkv + h * T5_D + 2 * c;
    const __nv_bfloat16* qrb = p.q + (row0 + qb_c) * p.ld_qkv + h * T5_D + 2 * c;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      aq[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(qra + ks * 16));
      aq[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(qrb + ks * 16));
      aq[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(qra + ks * 16 + 8));
      aq[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(qrb + ks * 16 + 8));
    }
  }
  // scores for every key tile
  float sc[T5M_NT][4];
  const __nv_bfloat16* ba = p.bias + (static_cast<long long>(h) * p.S + qa_c) * p.S;
  const __nv_bfloat16* bb = p.bias + (static_cast<long long>(h) * p.S + qb_c) * p.S;
  float mxa = -INFINITY, mxb = -INFINITY;
  const bool pair_loads = (p.S & 1) == 0 && (reinterpret_cast<uintptr_t>(p.bias) & 3) == 0;
#pragma unroll
  for (int nt = 0; nt < T5M_NT; ++nt) {
    sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
    if (nt < ntiles) {
      const __nv_bfloat16* kr = Ks + (nt * 8 + g) * T5M_LD + 2 * c;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_bf16_16816(sc[nt], aq[ks], *reinterpret_cast<const uint32_t*>(kr + ks * 16), *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8));
      const int j = nt * 8 + 2 * c;
      float bias_a[2], bias_b[2];
      if (pair_loads && j + 1 < p.S) {                                         // even S: the (j, j+1) pair of a bias row is 4-byte aligned
        const float2 fa = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(ba + j)));
        const float2 fb = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(bb + j)));
        bias_a[0] = fa.x; bias_a[1] = fa.y; bias_b[0] = fb.x; bias_b[1] = fb.y;
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jj = (j + e < p.S) ? j + e : 0;
          bias_a[e] = __bfloat162float(__ldg(ba + jj));
          bias_b[e] = __bfloat162float(__ldg(bb + jj));
        }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = j + e < p.S;
        // matmul output rounded to bf16, then `scores += position_bias` rounded again
        const float va = bf16_round(bf16_round(sc[nt][e]) + bias_a[e]);
        const float vb = bf16_round(bf16_round(sc[nt][2 + e]) + bias_b[e]);
        sc[nt][e] = ok ? va : -INFINITY;
        sc[nt][2 + e] = ok ? vb : -INFINITY;
      }
      mxa = fmaxf(mxa, fmaxf(sc[nt][0], sc[nt][1]));
      mxb = fmaxf(mxb, fmaxf(sc[nt][2], sc[nt][3]));
    }
  }
  mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 1)); mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 2));
  mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 1)); mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 2));
  float suma = 0.f, sumb = 0.f;
#pragma unroll
  for (int nt = 0; nt < T5M_NT; ++nt) {
    if (nt < ntiles) {
      sc[nt][0] = __expf(sc[nt][0] - mxa); sc[nt][1] = __expf(sc[nt][1] - mxa);
      sc[nt][2] = __expf(sc[nt][2] - mxb); sc[nt][3] = __expf(sc[nt][3] - mxb);
      suma += sc[nt][0] + sc[nt][1];
      sumb += sc[nt][2] + sc[nt][3];
    }
  }
  suma += __shfl_xor_sync(0xffffffffu, suma, 1); suma += __shfl_xor_sync(0xffffffffu, suma, 2);
  sumb += __shfl_xor_sync(0xffffffffu, sumb, 1); sumb += __shfl_xor_sync(0xffffffffu, sumb, 2);
  const float inva = 1.0f / suma, invb = 1.0f / sumb;
  // out = P V: the rounded probabilities of key tiles (2i, 2i+1) are the A fragment of k-step i
  float o[8][4];
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < T5M_NT / 2; ++kt) {
    if (2 * kt < ntiles) {                                                     // ntiles is even (Spad is a multiple of 16)
      uint32_t ap[4];
      ap[0] = pack_bf16x2(sc[2 * kt][0] * inva, sc[2 * kt][1] * inva);         // softmax(...).type_as(scores)
      ap[1] = pack_bf16x2(sc[2 * kt][2] * invb, sc[2 * kt][3] * invb);
      ap[2] = pack_bf16x2(sc[2 * kt + 1][0] * inva, sc[2 * kt + 1][1] * inva);
      ap[3] = pack_bf16x2(sc[2 * kt + 1][2] * invb, sc[2 * kt + 1][3] * invb);
      const unsigned short* vr = reinterpret_cast<const unsigned short*>(Vs + (kt * 16 + 2 * c) * T5M_LD + g);
#pragma unroll
      for (int dn = 0; dn < 8; ++dn) {
        // B fragment of V: rows = keys (2c, 2c+1) and (2c+8, 2c+9) of this k-step, column = dim dn*8 + g
        const uint32_t b0 = static_cast<uint32_t>(vr[dn * 8]) | (static_cast<uint32_t>(vr[dn * 8 + T5M_LD]) << 16);
        const uint32_t b1 = static_cast<uint32_t>(vr[dn * 8 + 8 * T5M_LD]) | (static_cast<uint32_t>(vr[dn * 8 + 9 * T5M_LD]) << 16);
        mma_bf16_16816(o[dn], ap, b0, b1);
      }
    }
  }
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) {
    if (qa < p.S) *reinterpret_cast<uint32_t*>(p.out + (row0 + qa) * p.ldo + h * T5_D + dn * 8 + 2 * c) = pack_bf16x2(o[dn][0], o[dn][1]);
    if (qb < p.S) *reinterpret_cast<uint32_t*>(p.out + (row0 + qb) * p.ldo + h * T5_D + dn * 8 + 2 * c) = pack_bf16x2(o[dn][2], o[dn][3]);
  }
}

__global__ void __launch_bounds__(256)
gated_mul_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out,
                 int rows, int nvec, long long lda, long long ldb, long long ldo) {
  const long long total = static_cast<long long>(rows) * nvec;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / nvec), c = static_cast<int>(i - static_cast<long long>(r) * nvec) * 8;
    const uint4 au = *reinterpret_cast<const uint4*>(a + r * lda + c);
    const uint4 bu = *reinterpret_cast<const uint4*>(b + r * ldb + c);
    const __nv_bfloat162* av = reinterpret_cast<const __nv_bfloat162*>(&au);
    const __nv_bfloat162* bv = reinterpret_cast<const __nv_bfloat162*>(&bu);
    uint4 ou;
    __nv_bfloat162* ov = reinterpret_cast<__nv_bfloat162*>(&ou);
#pragma unroll
    for (int k = 0; k < 4; ++k) ov[k] = __hmul2_rn(av[k], bv[k]);
    *reinterpret_cast<uint4*>(out + r * ldo + c) = ou;
  }
}

}  // namespace
}  // namespace vgpa

extern "C" int vgpa_t5_attention_bf16(const void* q, const void* k, const void* v, const void* bias, void* out, int B, int H,
                                      int S, int64_t ld_qkv, int64_t ldo, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(q && k && v && bias && out, "vgpa_t5_attention_bf16: null pointer");
  VGPA_CHECK(B > 0 && H > 0 && S > 0 && S <= T5_SMAX, "vgpa_t5_attention_bf16: need 0 < S <= %d (got B=%d H=%d S=%d)", T5_SMAX, B, H, S);
  VGPA_CHECK(ld_qkv % 8 == 0 && ld_qkv >= static_cast<int64_t>(H) * T5_D && ldo % 2 == 0 && ldo >= static_cast<int64_t>(H) * T5_D,
             "vgpa_t5_attention_bf16: leading dimensions must cover H * 64 columns (ld_qkv multiple of 8)");
  VGPA_CHECK(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(out) & 3) == 0, "vgpa_t5_attention_bf16: q/k/v must be 16-byte aligned");
  T5AttnParams p;
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ld_qkv = ld_qkv; p.ldo = ldo; p.S = S; p.H = H;
  static int use_mma = -1;
  if (use_mma < 0) { const char* e = getenv("VGPA_T5_ATTN_MMA"); use_mma = e ? atoi(e) : 1; }
  if (use_mma && S <= 8 * T5M_NT && ld_qkv % 2 == 0) {
    const int Spad = (S + 15) & ~15;
    const size_t smem_m = static_cast<size_t>(Spad) * T5M_LD * 2 * sizeof(__nv_bfloat16);
    static bool configured_m = false;
    if (!configured_m) {
      VGPA_CUDA(cudaFuncSetAttribute(t5_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
      configured_m = true;
    }
    const dim3 grid_m(static_cast<unsigned>(H), static_cast<unsigned>(B), static_cast<unsigned>((S + T5M_WARPS * 16 - 1) / (T5M_WARPS * 16)));
    t5_attention_mma_kernel<<<grid_m, T5M_WARPS * 32, smem_m, static_cast<cudaStream_t>(stream)>>>(p);
    VGPA_LAUNCH_CHECK("t5_attention_mma_kernel");
    return 0;
  }
  const size_t smem = static_cast<size_t>(S) * T5_PAD * 2 * sizeof(__nv_bfloat16) + static_cast<size_t>(T5_WARPS) * S * sizeof(float);
  static bool configured = false;
  if (!configured) {
    VGPA_CUDA(cudaFuncSetAttribute(t5_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  const dim3 grid(static_cast<unsigned>(H), static_cast<unsigned>(B), static_cast<unsigned>((S + T5_QB - 1) / T5_QB));
  t5_attention_kernel<<<grid, T5_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  VGPA_LAUNCH_CHECK("t5_attention_kernel");
  return 0;
}

extern "C" int vgpa_gated_mul_bf16(const void* a, const void* b, void* out, int rows, int N, int64_t lda, int64_t ldb,
                                   int64_t ldo, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a && b && out, "vgpa_gated_mul_bf16: null pointer");
  VGPA_CHECK(rows > 0 && N > 0 && N % 8 == 0, "vgpa_gated_mul_bf16: N=%d must be a positive multiple of 8", N);
  VGPA_CHECK(lda >= N && ldb >= N && ldo >= N && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0, "vgpa_gated_mul_bf16: bad leading dimension");
  VGPA_CHECK(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             "vgpa_gated_mul_bf16: pointers must be 16-byte aligned");
  const long long total = static_cast<long long>(rows) * (N / 8);
  long long grid = (total + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  gated_mul_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(out), rows, N / 8,
      lda, ldb, ldo);
  VGPA_LAUNCH_CHECK("gated_mul_kernel");
  return 0;
}
