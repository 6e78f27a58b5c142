"""CogVideoX schedulers (host-side float64 tables + one fused CUDA update per step).

Mirrors diffusers' CogVideoXDDIMScheduler / CogVideoXDPMScheduler as the reference uses them
(generate/CogVideoX-5B.py:18 swaps in the DPM scheduler with timestep_spacing="trailing";
generate/CogVideoX-5B-I2V.py:18-19 keeps the checkpoint default; train/CogVideoX-5B/03_train.py:129-130,
154-155 use add_noise / get_velocity). Math: SURVEY.md App. A.4 — scaled_linear betas, zero-terminal-SNR
rescale, v-prediction. The per-element work (CFG combine + update) is vgpa_cfg_scheduler_step.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import dense


def cogvideox_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, snr_shift_scale=1.0) -> np.ndarray:
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float64) ** 2
    ac = np.cumprod(1.0 - betas)
    ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
    r = np.sqrt(ac)
    r0, rT = r[0], r[-1]
    r = (r - rT) * r0 / (r0 - rT)
    return r ** 2


class _Base:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, snr_shift_scale=1.0,
                 timestep_spacing="trailing"):
        if timestep_spacing != "trailing":
            raise RuntimeError("only timestep_spacing='trailing' (CogVideoX 5B) is implemented")
        self.num_train_timesteps = num_train_timesteps
        self.alphas_cumprod = cogvideox_alphas_cumprod(num_train_timesteps, beta_start, beta_end, snr_shift_scale)
        self.timesteps = None
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        n = self.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        self.timesteps = (np.round(np.arange(n, 0, -n / num_inference_steps)) - 1).astype(np.int64)
        self._x0_old = None
        self._t_old = None
        return self.timesteps

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _prev(self, t: int) -> int:
        return int(t) - self.num_train_timesteps // self.num_inference_steps

    # training-side helpers (03_train.py:129-130,154-155); plain fp32 tensor math on the caller's device
    def _sqrt_coeffs(self, x, timesteps):
        """sqrt(alpha_bar_t), sqrt(1 - alpha_bar_t) taken in float64 on the host table, then cast to x's dtype."""
        a = self.alphas_cumprod[np.asarray(timesteps.cpu())]
        shp = (-1,) + (1,) * (x.dim() - 1)
        sa = torch.as_tensor(np.sqrt(a), dtype=x.dtype, device=x.device).view(shp)
        sb = torch.as_tensor(np.sqrt(1.0 - a), dtype=x.dtype, device=x.device).view(shp)
        return sa, sb

    def add_noise(self, x, noise, timesteps):
        sa, sb = self._sqrt_coeffs(x, timesteps)
        return sa * x + sb * noise

    def get_velocity(self, x, noise, timesteps):
        sa, sb = self._sqrt_coeffs(x, timesteps)
        return sa * noise - sb * x


class CogVideoXDDIMScheduler(_Base):
    """Deterministic (eta = 0) DDIM; `step_cfg` fuses classifier-free guidance with the update."""

    def coefficients(self, t: int):
        ac = self.alphas_cumprod
        a_t = ac[int(t)]
        tp = self._prev(t)
        a_prev = ac[tp] if tp >= 0 else 1.0
        a = ((1 - a_prev) / (1 - a_t)) ** 0.5
        b = a_prev ** 0.5 - a_t ** 0.5 * a
        return dict(sqrt_alpha_t=a_t ** 0.5, sqrt_beta_t=(1 - a_t) ** 0.5, c_sample=a, c_x0=b)

    def step_cfg(self, pred_cond, pred_uncond, t: int, sample, guidance_scale: float, generator=None, out=None):
        k = self.coefficients(t)
        return dense.cfg_scheduler_step(pred_cond, pred_uncond, sample, mode=dense.SCHED_DDIM, guidance=guidance_scale,
                                        out=out, **k)


class CogVideoXDPMScheduler(_Base):
    """SDE DPM-Solver++ (2M) as in diffusers' CogVideoXDPMScheduler (App. A.4). Stochastic: consumes
    `generator` every step (one randn of the latent shape), in step order."""

    def _lam(self, a):
        # zero-terminal-SNR makes alphas_cumprod[999] exactly 0: lambda = -inf there (torch's log(0) in diffusers)
        return -math.inf if a <= 0.0 else math.log((a / (1 - a)) ** 0.5)

    def coefficients(self, t: int, t_back):
        ac = self.alphas_cumprod
        a_t = ac[int(t)]
        tp = self._prev(t)
        a_prev = ac[tp] if tp >= 0 else 1.0
        if a_prev >= 1.0:
            m1, m2, mn, h = 0.0, -1.0, 0.0, float("inf")
        else:
            h = self._lam(a_prev) - self._lam(a_t)
            m1 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * math.exp(-h)
            m2 = math.expm1(-2 * h) * a_prev ** 0.5
            mn = (1 - a_prev) ** 0.5 * (1 - math.exp(-2 * h)) ** 0.5
        c_x0, c_old = -m2, 0.0
        if t_back is not None and tp >= 0 and not math.isinf(h):
            r = (self._lam(a_t) - self._lam(ac[int(t_back)])) / h
            c_x0 = -m2 * (1 + 1 / (2 * r))
            c_old = m2 / (2 * r)
        return dict(sqrt_alpha_t=a_t ** 0.5, sqrt_beta_t=(1 - a_t) ** 0.5, c_sample=m1, c_x0=c_x0, c_x0_old=c_old, c_noise=mn)

    def step_cfg(self, pred_cond, pred_uncond, t: int, sample, guidance_scale: float, generator=None, out=None):
        k = self.coefficients(t, self._t_old)
        noise = torch.randn(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
        x0_out = torch.empty(sample.shape, dtype=torch.float32, device=sample.device)
        prev = dense.cfg_scheduler_step(pred_cond, pred_uncond, sample, mode=dense.SCHED_DPM, guidance=guidance_scale,
                                        x0_old=self._x0_old if k["c_x0_old"] != 0.0 else None, x0_out=x0_out, noise=noise,
                                        out=out, **k)
        self._x0_old, self._t_old = x0_out, int(t)
        return prev
