// K9: DPO preference loss over six velocity tensors — one streaming pass forward, one backward.
//
// Replaces DPOLoss.forward (train/loss.py:53-121 of the reference; called from
// train/CogVideoX-5B/03_train.py:157): four per-sample MSEs, logits = beta*((ref_w-mod_w)-(ref_l-mod_l)),
// -logsigmoid / BCE-with-smoothing / hinge, plus the logged statistics. The reference spends 12+
// elementwise/reduce launches and re-reads each target twice; here every element of the six
// tensors is read exactly once (HBM-bound, 128-bit loads), partial sums are fp64 per block and the
// final reduction runs in a fixed order (deterministic). Each tensor is fp32 or bf16 independently
// (under Lightning bf16-mixed the predictions are bf16 and the targets fp32); math is fp32.
#include "common.cuh"
#include "../../include/videogpa_b200.h"
#include <math.h>

namespace vgpa {
namespace {

constexpr int DL_THREADS = 256;

template <bool BF>
__device__ __forceinline__ void load4(const void* p, long long i, float (&o)[4]) {
  if (BF) {
    const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p) + i);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  } else {
    const float4 f = *reinterpret_cast<const float4*>(static_cast<const float*>(p) + i);
    o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = f.w;
  }
}
__device__ __forceinline__ void load4_dyn(const void* p, int is_bf16, long long i, float (&o)[4]) {
  if (is_bf16) load4<true>(p, i, o); else load4<false>(p, i, o);
}
__device__ __forceinline__ float load1_dyn(const void* p, int is_bf16, long long i) {
  return is_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]) : static_cast<const float*>(p)[i];
}

struct DlTensors {
  const void* p[6];   // v_win, v_lose, v_win_ref, v_lose_ref, v_win_target, v_lose_target
  int bf[6];
};

__global__ void __launch_bounds__(DL_THREADS)
dpo_partial_kernel(DlTensors t, long long n, int blocks_per_sample, double* __restrict__ partial) {
  const int b = blockIdx.y;
  const long long off = static_cast<long long>(b) * n;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // model_win, model_lose, ref_win, ref_lose
  double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
  const long long n4 = n & ~3LL;
  const bool aligned = (off & 3) == 0;
  int flush = 0;
  if (aligned) {
    for (long long i = (static_cast<long long>(blockIdx.x) * DL_THREADS + threadIdx.x) * 4; i < n4;
         i += static_cast<long long>(blocks_per_sample) * DL_THREADS * 4) {
      float vw[4], vl[4], rw[4], rl[4], tw[4], tl[4];
      load4_dyn(t.p[0], t.bf[0], off + i, vw); load4_dyn(t.p[1], t.bf[1], off + i, vl);
      load4_dyn(t.p[2], t.bf[2], off + i, rw); load4_dyn(t.p[3], t.bf[3], off + i, rl);
      load4_dyn(t.p[4], t.bf[4], off + i, tw); load4_dyn(t.p[5], t.bf[5], off + i, tl);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = vw[k] - tw[k], c = vl[k] - tl[k], e = rw[k] - tw[k], g = rl[k] - tl[k];
        s0 = fmaf(a, a, s0); s1 = fmaf(c, c, s1); s2 = fmaf(e, e, s2); s3 = fmaf(g, g, s3);
      }
      if (++flush == 64) { d0 += s0; d1 += s1; d2 += s2; d3 += s3; s0 = s1 = s2 = s3 = 0.f; flush = 0; }
    }
  }
  // scalar tail (or the whole sample when its offset is not 4-aligned)
  for (long long i = (aligned ? n4 : 0) + static_cast<long long>(blockIdx.x) * DL_THREADS + threadIdx.x; i < n;
       i += static_cast<long long>(blocks_per_sample) * DL_THREADS) {
    const float tw = load1_dyn(t.p[4], t.bf[4], off + i), tl = load1_dyn(t.p[5], t.bf[5], off + i);
    const float a = load1_dyn(t.p[0], t.bf[0], off + i) - tw, c = load1_dyn(t.p[1], t.bf[1], off + i) - tl;
    const float e = load1_dyn(t.p[2], t.bf[2], off + i) - tw, g = load1_dyn(t.p[3], t.bf[3], off + i) - tl;
    d0 += static_cast<double>(a * a); d1 += static_cast<double>(c * c);
    d2 += static_cast<double>(e * e); d3 += static_cast<double>(g * g);
  }
  d0 += s0; d1 += s1; d2 += s2; d3 += s3;
  d0 = warp_sum_d(d0); d1 = warp_sum_d(d1); d2 = warp_sum_d(d2); d3 = warp_sum_d(d3);
  __shared__ double sh[DL_THREADS / 32][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[warp][0] = d0; sh[warp][1] = d1; sh[warp][2] = d2; sh[warp][3] = d3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double tot = 0.0;
    for (int w = 0; w < DL_THREADS / 32; ++w) tot += sh[w][threadIdx.x];
    partial[(static_cast<long long>(b) * blocks_per_sample + blockIdx.x) * 4 + threadIdx.x] = tot;
  }
}

__device__ __forceinline__ float softplus_f(float x) {  // log(1 + exp(x)), stable
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

// out5: loss, reward_margin, winner_reward, loser_reward, accuracy. err4: [4][B]. coef: [2][B] = dL/d(model_{win,lose}_err)
__global__ void dpo_finalize_kernel(const double* __restrict__ partial, int B, long long n, int blocks_per_sample,
                                    float beta, float label_smoothing, int loss_type, float* __restrict__ out5,
                                    float* __restrict__ err4, float* __restrict__ coef) {
  // one warp: lane l sums the block partials l, l + 32, ... in order, then a fixed shuffle tree (deterministic; a single
  // thread walking several hundred dependent loads cost 17 us, more than the pass over the six tensors)
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  float loss = 0.f, margin = 0.f, wr = 0.f, lr = 0.f, acc = 0.f;
  for (int b = 0; b < B; ++b) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = lane; k < blocks_per_sample; k += 32)
      for (int c = 0; c < 4; ++c) s[c] += partial[(static_cast<long long>(b) * blocks_per_sample + k) * 4 + c];
    for (int c = 0; c < 4; ++c) s[c] = warp_sum_d(s[c]);
    const float mw = static_cast<float>(s[0] / static_cast<double>(n)), ml = static_cast<float>(s[1] / static_cast<double>(n));
    const float rw = static_cast<float>(s[2] / static_cast<double>(n)), rl = static_cast<float>(s[3] / static_cast<double>(n));
    if (err4 && lane == 0) { err4[0 * B + b] = mw; err4[1 * B + b] = ml; err4[2 * B + b] = rw; err4[3 * B + b] = rl; }
    const float win_diff = rw - mw, lose_diff = rl - ml;                 // loss.py:82-83
    const float logit = beta * (win_diff - lose_diff);                   // loss.py:93
    float l, dl;
    if (loss_type == VGPA_DPO_SIGMOID) {
      if (label_smoothing > 0.f) {                                       // BCE-with-logits, target 1 - eps (:97-103)
        const float tgt = 1.0f - label_smoothing;
        l = (1.0f - tgt) * logit + softplus_f(-logit);
        dl = 1.0f / (1.0f + expf(-logit)) - tgt;
      } else {                                                           // -logsigmoid (:105)
        l = softplus_f(-logit);
        dl = -1.0f / (1.0f + expf(logit));
      }
    } else if (loss_type == VGPA_DPO_HINGE) {                            // hinge (:106-108)
      l = fmaxf(1.0f - logit, 0.f);
      dl = (1.0f - logit) > 0.f ? -1.0f : 0.f;
    } else {                                                             // "sft": F.mse_loss(v_pred, v_target) (:141-143)
      l = mw;
      dl = 0.f;
    }
    loss += l;
    const float w_rew = -mw, l_rew = -ml;                                // loss.py:86-88
    margin += w_rew - l_rew; wr += w_rew; lr += l_rew;
    acc += (w_rew > l_rew) ? 1.0f : 0.f;
    if (coef && lane == 0) {
      // dlogit/d(model_win_err) = -beta, dlogit/d(model_lose_err) = +beta; mean over B
      coef[0 * B + b] = (loss_type == VGPA_DPO_SFT) ? 1.0f / static_cast<float>(B) : dl * (-beta) / static_cast<float>(B);
      coef[1 * B + b] = (loss_type == VGPA_DPO_SFT) ? 0.f : dl * (beta) / static_cast<float>(B);
    }
  }
  const float inv = 1.0f / static_cast<float>(B);
  if (lane != 0) return;
  out5[0] = loss * inv; out5[1] = margin * inv; out5[2] = wr * inv; out5[3] = lr * inv; out5[4] = acc * inv;
}

// grad_v[b, i] = grad_loss * coef[b] * 2 * (v - target) / n
__global__ void __launch_bounds__(DL_THREADS)
dpo_backward_kernel(const void* v_win, int bf_vw, const void* v_lose, int bf_vl, const void* t_win, int bf_tw,
                    const void* t_lose, int bf_tl, const float* __restrict__ coef, const float* __restrict__ grad_loss,
                    int B, long long n, void* g_win, void* g_lose) {
  const int b = blockIdx.y;
  const long long off = static_cast<long long>(b) * n;
  const float gl = grad_loss ? *grad_loss : 1.0f;
  const float cw = gl * coef[b] * 2.0f / static_cast<float>(n);
  const float cl = gl * coef[B + b] * 2.0f / static_cast<float>(n);
  for (long long i = static_cast<long long>(blockIdx.x) * DL_THREADS + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * DL_THREADS) {
    const float gw = cw * (load1_dyn(v_win, bf_vw, off + i) - load1_dyn(t_win, bf_tw, off + i));
    const float gg = cl * (load1_dyn(v_lose, bf_vl, off + i) - load1_dyn(t_lose, bf_tl, off + i));
    if (bf_vw) static_cast<__nv_bfloat16*>(g_win)[off + i] = __float2bfloat16_rn(gw);
    else static_cast<float*>(g_win)[off + i] = gw;
    if (bf_vl) static_cast<__nv_bfloat16*>(g_lose)[off + i] = __float2bfloat16_rn(gg);
    else static_cast<float*>(g_lose)[off + i] = gg;
  }
}

int dl_blocks_per_sample(int B, long long n) {
  long long need = (n + DL_THREADS * 4 * 2 - 1) / (DL_THREADS * 4 * 2);   // >= 2 vector iterations per thread: one pair (1.1 M elements) fills 148 SMs
  long long want = (148LL * 8 + B - 1) / B;
  if (need > want) need = want;
  if (need < 1) need = 1;
  return static_cast<int>(need);
}

}  // namespace
}  // namespace vgpa

extern "C" size_t vgpa_dpo_workspace_bytes(int B, int64_t n_per_sample) {
  if (B <= 0 || n_per_sample <= 0) return 256;
  return static_cast<size_t>(B) * vgpa::dl_blocks_per_sample(B, n_per_sample) * 32 + 256;
}

extern "C" int vgpa_dpo_loss_forward(const vgpa_dpo_args* a, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr, "vgpa_dpo_loss_forward: null args");
  VGPA_CHECK(a->B > 0 && a->n_per_sample > 0, "vgpa_dpo_loss_forward: bad shape B=%d n=%lld", a->B, (long long)a->n_per_sample);
  VGPA_CHECK(a->loss_type == VGPA_DPO_SIGMOID || a->loss_type == VGPA_DPO_HINGE || a->loss_type == VGPA_DPO_SFT, "vgpa_dpo_loss_forward: unknown loss type %d", a->loss_type);
  VGPA_CHECK(a->d_out5 && a->d_workspace, "vgpa_dpo_loss_forward: null output / workspace");
  VGPA_CHECK(a->workspace_bytes >= vgpa_dpo_workspace_bytes(a->B, a->n_per_sample), "vgpa_dpo_loss_forward: workspace too small");
  DlTensors t;
  for (int i = 0; i < 6; ++i) {
    VGPA_CHECK(a->tensors[i] != nullptr, "vgpa_dpo_loss_forward: tensor %d is null", i);
    VGPA_CHECK(a->is_bf16[i] == 0 || a->is_bf16[i] == 1, "vgpa_dpo_loss_forward: dtype flag %d invalid", i);
    VGPA_CHECK((reinterpret_cast<uintptr_t>(a->tensors[i]) & 15) == 0, "vgpa_dpo_loss_forward: tensor %d must be 16-byte aligned", i);
    t.p[i] = a->tensors[i];
    t.bf[i] = a->is_bf16[i];
  }
  VGPA_CHECK(a->B <= 65535, "vgpa_dpo_loss_forward: B too large");
  const int bps = dl_blocks_per_sample(a->B, a->n_per_sample);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(a->d_workspace);
  dpo_partial_kernel<<<dim3(bps, a->B), DL_THREADS, 0, s>>>(t, a->n_per_sample, bps, partial);
  VGPA_LAUNCH_CHECK("dpo_partial_kernel");
  dpo_finalize_kernel<<<1, 32, 0, s>>>(partial, a->B, a->n_per_sample, bps, a->beta, a->label_smoothing, a->loss_type,
                                      a->d_out5, a->d_err4, a->d_coef);
  VGPA_LAUNCH_CHECK("dpo_finalize_kernel");
  return 0;
}

extern "C" int vgpa_dpo_loss_backward(const vgpa_dpo_args* a, const float* d_grad_loss, void* d_grad_win, void* d_grad_lose,
                                      void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr && a->d_coef != nullptr, "vgpa_dpo_loss_backward: needs the coefficient buffer written by forward");
  VGPA_CHECK(d_grad_win && d_grad_lose, "vgpa_dpo_loss_backward: null gradient pointer");
  VGPA_CHECK(a->B > 0 && a->B <= 65535 && a->n_per_sample > 0, "vgpa_dpo_loss_backward: bad shape");
  long long bx = (a->n_per_sample + DL_THREADS * 8 - 1) / (DL_THREADS * 8);
  if (bx > 4096) bx = 4096;
  if (bx < 1) bx = 1;
  dpo_backward_kernel<<<dim3(static_cast<unsigned>(bx), a->B), DL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      a->tensors[0], a->is_bf16[0], a->tensors[1], a->is_bf16[1], a->tensors[4], a->is_bf16[4], a->tensors[5], a->is_bf16[5],
      a->d_coef, d_grad_loss, a->B, a->n_per_sample, d_grad_win, d_grad_lose);
  VGPA_LAUNCH_CHECK("dpo_backward_kernel");
  return 0;
}
