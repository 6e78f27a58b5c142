// Classifier-free guidance + one v-prediction scheduler update, fused into a single pass over the
// latent (diffusers CogVideoXPipeline.__call__ + CogVideoXDDIMScheduler.step /
// CogVideoXDPMScheduler.step, reached from generate/CogVideoX-5B.py:72-77).
// fp32 math, coefficients come from the host-side float64 schedule tables. Rounding points follow
// eager torch type promotion: `coef * sample` with a bf16 sample stays bf16, everything touching
// the fp32 noise prediction is fp32, the result is cast to bf16 once.
#include "common.cuh"
#include "../../include/videogpa_b200.h"

namespace vgpa {
namespace {

__global__ void __launch_bounds__(256)
cfg_step_kernel(vgpa_sched_args a) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const __nv_bfloat16* u = static_cast<const __nv_bfloat16*>(a.pred_uncond);
  const __nv_bfloat16* c = static_cast<const __nv_bfloat16*>(a.pred_cond);
  const __nv_bfloat16* xs = static_cast<const __nv_bfloat16*>(a.sample);
  float v = __bfloat162float(c[i]);
  if (u != nullptr) {
    const float uf = __bfloat162float(u[i]);
    v = uf + a.guidance * (v - uf);
  }
  const float x = __bfloat162float(xs[i]);
  const float x0 = bf16_round(a.sqrt_alpha_t * x) - a.sqrt_beta_t * v;
  float prev;
  if (a.mode == VGPA_SCHED_DDIM) {
    prev = bf16_round(a.c_sample * x) + a.c_x0 * x0;
  } else {
    float d = x0;
    if (a.x0_old != nullptr && a.c_x0_old != 0.f) d = a.c_x0 * x0 + a.c_x0_old * a.x0_old[i];
    else d = a.c_x0 * x0;
    prev = bf16_round(a.c_sample * x) + d;
    if (a.noise != nullptr) prev += a.c_noise * __bfloat162float(static_cast<const __nv_bfloat16*>(a.noise)[i]);
  }
  if (a.x0_out != nullptr) a.x0_out[i] = x0;
  static_cast<__nv_bfloat16*>(a.prev_sample)[i] = __float2bfloat16_rn(prev);
}

}  // namespace
}  // namespace vgpa

extern "C" int vgpa_cfg_scheduler_step(const vgpa_sched_args* a, void* stream) {
  using namespace vgpa;
  VGPA_CHECK(a != nullptr, "vgpa_cfg_scheduler_step: null args");
  VGPA_CHECK(a->n > 0 && a->pred_cond && a->sample && a->prev_sample, "vgpa_cfg_scheduler_step: null tensor / empty");
  VGPA_CHECK(a->mode == VGPA_SCHED_DDIM || a->mode == VGPA_SCHED_DPM, "vgpa_cfg_scheduler_step: unknown mode %d", a->mode);
  const unsigned grid = static_cast<unsigned>((a->n + 255) / 256);
  cfg_step_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  VGPA_LAUNCH_CHECK("cfg_step_kernel");
  return 0;
}
