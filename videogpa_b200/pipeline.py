"""Denoising loop of the CogVideoX pipelines on the sm_100a kernels.

Mirrors the part of diffusers' CogVideoXPipeline / CogVideoXImageToVideoPipeline.__call__ that the
reference drives (generate/CogVideoX-5B.py:72-77, generate/CogVideoX-5B-I2V.py:79-88, SURVEY.md §3.1 and
App. A.4): latents = randn * init_noise_sigma; per step the CFG pair [uncond; cond] goes through the
transformer, guidance is combined in fp32 and the scheduler updates the latent. Text encoding (T5-XXL)
and VAE weights are third-party checkpoints that are not reachable here, so the pipeline takes
`prompt_embeds` / `negative_prompt_embeds` (as diffusers pipelines also accept) and returns latents
unless a decoder is supplied.

`cfg_group` shards the cond/uncond pair over two ranks: each rank runs one branch and
the two noise predictions are exchanged with one NCCL all-gather per step; both ranks then apply the
same update so their latents stay bit-identical without a broadcast.
"""
from __future__ import annotations

import math

import torch

from .rope import get_3d_rotary_pos_embed

BF16 = torch.bfloat16


class CogVideoXDenoisePipeline:
    vae_scale_factor_spatial = 8
    vae_scale_factor_temporal = 4

    def __init__(self, transformer, scheduler, vae=None, vae_scaling_factor: float = 0.7):
        self.transformer = transformer
        self.scheduler = scheduler
        self.vae = vae
        self.vae_scaling_factor = vae_scaling_factor
        self.device = transformer.device
        # pinned staging buffers for the host-facing step (allocated on first use)
        self._pin = {}

    # ------------------------------------------------------------------ helpers
    def latent_shape(self, batch: int, num_frames: int, height: int, width: int):
        c = self.transformer.config
        return (batch, (num_frames - 1) // self.vae_scale_factor_temporal + 1, c.out_channels,
                height // self.vae_scale_factor_spatial, width // self.vae_scale_factor_spatial)

    def prepare_latents(self, batch, num_frames, height, width, generator=None):
        shape = self.latent_shape(batch, num_frames, height, width)
        lat = torch.randn(shape, generator=generator, device=self.device, dtype=BF16)
        return lat * self.scheduler.init_noise_sigma

    def rotary(self, latent_frames: int, latent_h: int, latent_w: int):
        c = self.transformer.config
        if not c.use_rotary_positional_embeddings:
            return None
        pt = getattr(c, "patch_size_t", None) or 1                      # 1.5: one rotary position per temporal patch
        return get_3d_rotary_pos_embed(c.attention_head_dim, latent_h // c.patch_size, latent_w // c.patch_size,
                                       (latent_frames + pt - 1) // pt, device=self.device)

    @staticmethod
    def dynamic_guidance(guidance_scale: float, num_inference_steps: int, t: int) -> float:
        """use_dynamic_cfg of the 1.5 script (generate/CogVideoX1.5-5B.py:85; App. A.4, raw timestep value, sic)."""
        return 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_inference_steps - t) / num_inference_steps) ** 5.0)) / 2)

    # ------------------------------------------------------------------ one denoise step, device resident
    def denoise_step(self, latents, prompt_embeds_pair, t: int, guidance_scale: float, rope, image_latents=None,
                     generator=None, cfg_group=None, out=None):
        """latents [B,F,C,H,W] bf16; prompt_embeds_pair [2B,St,4096] = [uncond; cond] -> next latents."""
        B = latents.shape[0]
        if cfg_group is None:
            x = torch.cat([latents, latents], dim=0)
            if image_latents is not None:
                x = torch.cat([x, torch.cat([image_latents, image_latents], dim=0)], dim=2)
            ts = torch.full((2 * B,), float(t), device=self.device, dtype=torch.float32)
            pred = self.transformer(hidden_states=x, encoder_hidden_states=prompt_embeds_pair, timestep=ts,
                                    image_rotary_emb=rope, return_dict=False)[0]
            pred_uncond, pred_cond = pred[:B], pred[B:]
        else:
            x = latents if image_latents is None else torch.cat([latents, image_latents], dim=2)
            mine = prompt_embeds_pair[cfg_group.branch * B:(cfg_group.branch + 1) * B]
            ts = torch.full((B,), float(t), device=self.device, dtype=torch.float32)
            pred = self.transformer(hidden_states=x, encoder_hidden_states=mine, timestep=ts, image_rotary_emb=rope,
                                    return_dict=False)[0]
            if hasattr(cfg_group, "peer_views"):
                # exchange fused into the guidance + scheduler kernel: the partner's prediction is read over NVLink peer memory
                pred_uncond, pred_cond = cfg_group.peer_views(pred.contiguous())
                nxt = self.scheduler.step_cfg(pred_cond, pred_uncond, int(t), latents, guidance_scale, generator=generator, out=out)
                cfg_group.release()
                return nxt
            pred_uncond, pred_cond = cfg_group.exchange(pred)
        return self.scheduler.step_cfg(pred_cond.contiguous(), pred_uncond.contiguous(), int(t), latents, guidance_scale,
                                       generator=generator, out=out)

    # ------------------------------------------------------------------ one denoise step, host buffers in / out
    def denoise_step_host(self, latents_host: torch.Tensor, prompt_embeds_pair_host: torch.Tensor, t: int,
                          guidance_scale: float, rope, out_host: torch.Tensor | None = None) -> torch.Tensor:
        """Public host-facing call: pinned host latents + prompt embeddings in, next latents on the host out.
        The H2D / D2H copies run on the current stream around the same kernels as `denoise_step`."""
        lat = latents_host.to(self.device, non_blocking=True)
        pe = prompt_embeds_pair_host.to(self.device, non_blocking=True)
        nxt = self.denoise_step(lat, pe, t, guidance_scale, rope)
        if out_host is None:
            out_host = torch.empty(nxt.shape, dtype=nxt.dtype, pin_memory=True)
        out_host.copy_(nxt, non_blocking=True)
        return out_host

    # ------------------------------------------------------------------ the full loop
    @torch.no_grad()
    def __call__(self, prompt_embeds, negative_prompt_embeds, *, num_frames=49, height=480, width=720, num_inference_steps=50,
                 guidance_scale=6.0, use_dynamic_cfg=False, generator=None, latents=None, image_latents=None,
                 cfg_group=None, output_type="latent", callback=None):
        pe = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0).to(device=self.device, dtype=BF16).contiguous()
        B = prompt_embeds.shape[0]
        timesteps = self.scheduler.set_timesteps(num_inference_steps)
        # CogVideoX1.5 (patch_size_t = 2): the latent frame count is padded to a multiple of the temporal patch by generating
        # extra leading frames, which are dropped before decoding (diffusers CogVideoXPipeline.__call__, App. A.4)
        pt = getattr(self.transformer.config, "patch_size_t", None) or 1
        latent_frames = (num_frames - 1) // self.vae_scale_factor_temporal + 1
        additional_frames = (pt - latent_frames % pt) % pt
        if latents is None:
            latents = self.prepare_latents(B, num_frames + additional_frames * self.vae_scale_factor_temporal, height, width, generator)
        latents = latents.to(device=self.device, dtype=BF16).contiguous()
        rope = self.rotary(latents.shape[1], latents.shape[3], latents.shape[4])
        if image_latents is not None:
            image_latents = image_latents.to(device=self.device, dtype=BF16).contiguous()
        for i, t in enumerate(timesteps):
            g = self.dynamic_guidance(guidance_scale, num_inference_steps, int(t)) if use_dynamic_cfg else guidance_scale
            latents = self.denoise_step(latents, pe, int(t), g, rope, image_latents=image_latents, generator=generator,
                                        cfg_group=cfg_group)
            if callback is not None:
                callback(i, int(t), latents)
        if additional_frames:
            latents = latents[:, additional_frames:]
        if output_type == "latent" or self.vae is None:
            return latents
        # decode: latents [B,F,C,H,W] -> [B,C,F,H,W] / scaling_factor (App. A.4)
        return self.vae.decode(latents.permute(0, 2, 1, 3, 4) / self.vae_scaling_factor).sample
