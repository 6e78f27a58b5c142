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

    _CONFIG_KEYS = ("num_train_timesteps", "beta_start", "beta_end", "snr_shift_scale", "timestep_spacing")

    @classmethod
    def from_config(cls, config: dict, **overrides):
        """diffusers' `Scheduler.from_config(pipe.scheduler.config, timestep_spacing="trailing")` (generate/CogVideoX-5B.py:18):
        the checkpoint's betas / snr_shift_scale decide alphas_cumprod (CogVideoX-2B ships snr_shift_scale 3.0, 5B 1.0).
        Settings this implementation cannot honour raise instead of silently producing other noise levels."""
        cfg = dict(config)
        cfg.update(overrides)
        if cfg.get("beta_schedule", "scaled_linear") != "scaled_linear":
            raise RuntimeError(f"beta_schedule {cfg['beta_schedule']!r} is not implemented (CogVideoX uses scaled_linear)")
        if cfg.get("prediction_type", "v_prediction") != "v_prediction":
            raise RuntimeError(f"prediction_type {cfg['prediction_type']!r} is not implemented (CogVideoX uses v_prediction)")
        if not cfg.get("rescale_betas_zero_snr", True):
            raise RuntimeError("rescale_betas_zero_snr = false is not implemented (CogVideoX checkpoints set it)")
        return cls(**{k: cfg[k] for k in cls._CONFIG_KEYS if k in cfg})

    @classmethod
    def from_pretrained(cls, model_dir, subfolder: str = "scheduler", **overrides):
        """`from_pretrained(base, subfolder="scheduler")` (train/CogVideoX-5B/03_train.py:96-113): reads
        <model_dir>/<subfolder>/scheduler_config.json."""
        import json
        import os
        path = os.path.join(str(model_dir), subfolder, "scheduler_config.json")
        if not os.path.isfile(path):
            raise RuntimeError(f"{path} not found")
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_config(json.load(f), **overrides)

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


class FlowUniPCMultistepScheduler:
    """The sampler `WanTI2V.generate` uses by default (`sample_solver='unipc'`; Wan2.2 `wan/utils/fm_solvers_unipc.py`, a
    flow-matching variant of diffusers' UniPCMultistepScheduler): UniPC with B(h) = expm1(h) ("bh2"), data prediction,
    solver_order 2, lower_order_final, corrector on every step after the first. Reference call site:
    generate/Wan2.2-TI2V-5B.py:120-129 (`engine.generate(..., shift, sampling_steps)`). The un-vendored Wan2.2 repository is
    not in /root/reference, so this follows the published algorithm: **parity unpinned**; tests check it against the exact
    solution of a linear flow (convergence order) and against Euler.

    Host-side coefficient arithmetic in float64, tensor updates as fp32 linear combinations (the latent is 14 MB: the
    sampler update is not a hot op; the two DiT forwards per step are)."""

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, shift: float = 1.0, lower_order_final: bool = True):
        self.num_train_timesteps, self.solver_order, self.lower_order_final = num_train_timesteps, solver_order, lower_order_final
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sig = 1.0 - alphas
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.sigma_min, self.sigma_max = float(sig[-1]), float(sig[0])
        self.sigmas = None

    def set_timesteps(self, num_inference_steps: int, shift: float = 1.0, sigmas=None):
        """sigmas (optional): an explicit decreasing grid of num_inference_steps + 1 values instead of the shifted linear one."""
        if sigmas is not None:
            full = np.asarray(sigmas, dtype=np.float64)
            if full.shape != (num_inference_steps + 1,) or np.any(np.diff(full) >= 0):
                raise RuntimeError("sigmas must be a strictly decreasing grid of num_inference_steps + 1 values")
            sig = full[:-1]
        else:
            sig = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
            sig = shift * sig / (1 + (shift - 1) * sig)
            full = np.concatenate([sig, [0.0]])
        self.timesteps = (sig * self.num_train_timesteps).tolist()
        self.sigmas = full.astype(np.float64)
        self.num_inference_steps = num_inference_steps
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = 0
        self.this_order = 1

    @staticmethod
    def _lam(sigma: float) -> float:
        alpha = 1.0 - sigma
        if sigma <= 0.0:
            return math.inf
        if alpha <= 0.0:
            return -math.inf
        return math.log(alpha) - math.log(sigma)

    def _rhos(self, order: int, hh: float, rks: list, corrector: bool):
        """rhos of the UniP / UniC update for B(h) = expm1(h)."""
        h_phi_1 = math.expm1(hh) if math.isfinite(hh) else (-1.0 if hh < 0 else math.inf)
        B_h = h_phi_1
        h_phi_k = (h_phi_1 / hh - 1.0) if math.isfinite(hh) else -1.0
        R, b = [], []
        fact = 1
        rk = np.asarray(rks, dtype=np.float64)
        for i in range(1, order + 1):
            R.append(rk ** (i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = (h_phi_k / hh - 1.0 / fact) if math.isfinite(hh) else (-1.0 / fact)
        R, b = np.stack(R), np.asarray(b)
        if corrector:
            rhos = np.array([0.5]) if order == 1 else np.linalg.solve(R, b)
        else:
            rhos = np.array([0.5]) if order == 2 else (np.linalg.solve(R[:-1, :-1], b[:-1]) if order > 2 else np.array([]))
        return h_phi_1, B_h, rhos

    def _update(self, x, m0, sigma_s0: float, sigma_t: float, order: int, base_index: int, model_t=None):
        """Common part of UniP (model_t None) and UniC: x_t from x (at sigma_s0) with the stored data predictions."""
        alpha_t = 1.0 - sigma_t
        lam_t, lam_s0 = self._lam(sigma_t), self._lam(sigma_s0)
        h = lam_t - lam_s0
        rks, D1s = [], []
        for i in range(1, order):
            mi = self.model_outputs[-(i + 1)]
            rk = (self._lam(float(self.sigmas[base_index - i])) - lam_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        h_phi_1, B_h, rhos = self._rhos(order, -h, rks, corrector=model_t is not None)
        x_t = (sigma_t / sigma_s0) * x - (alpha_t * h_phi_1) * m0
        res = None
        for k, D in enumerate(D1s):
            term = float(rhos[k]) * D
            res = term if res is None else res + term
        if model_t is not None:
            term = float(rhos[-1]) * (model_t - m0)
            res = term if res is None else res + term
        if res is not None:
            x_t = x_t - (alpha_t * B_h) * res
        return x_t

    def step(self, model_output: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        """model_output = predicted velocity at (sample, sigmas[step_index]) -> sample at sigmas[step_index + 1] (fp32)."""
        if self.sigmas is None:
            raise RuntimeError("call set_timesteps first")
        i = self.step_index
        sample = sample.float()
        sigma = float(self.sigmas[i])
        x0 = sample - sigma * model_output.float()                        # flow matching: x0 = x_t - sigma_t * v
        if i > 0 and self.last_sample is not None:                        # UniC: correct the sample with the new prediction
            sample = self._update(self.last_sample, self.model_outputs[-1], float(self.sigmas[i - 1]), sigma, self.this_order,
                                  base_index=i - 1, model_t=x0)
        self.model_outputs = self.model_outputs[1:] + [x0]
        order = min(self.solver_order, self.num_inference_steps - i) if self.lower_order_final else self.solver_order
        self.this_order = min(order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev = self._update(sample, x0, sigma, float(self.sigmas[i + 1]), self.this_order, base_index=i)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev
