"""Forward half of the DPO `_shared_step` (reference train/CogVideoX-5B/03_train.py:116-157) — what the
reference's `validation_step` (:189-201) runs under `torch.no_grad()`:

    x_win, x_lose [B, C, F, H, W] -> permute -> [B, F, C, H, W];  t ~ U{0..999};  shared noise
    x_*_noisy = scheduler.add_noise(x_*, noise, t)
    v_*_pred = transformer(x_*_noisy, prompt_emb, t)   (policy = base + merged LoRA)   NO image_rotary_emb (the quirk of :134-139)
    v_*_ref  = ref_transformer(x_*_noisy, prompt_emb, t)
    v_*_target = scheduler.get_velocity(x_*, noise, t)
    loss_fn(v_win_pred, v_lose_pred, v_win_ref, v_lose_ref, v_win_target, v_lose_target) -> LossOutput

The winner and the loser share noise, timestep and prompt, so each model sees them as one batch of 2B samples (one pass over
its weights). The four forwards run on the sm_100a DiT kernels and the loss on the fused DPO kernel. `training_step` runs the
policy through videogpa_b200.train_dit (differentiable forward, hand-written backward kernels, LoRA factors as the only
trainable parameters) and returns the loss for `backward()`.
"""
from __future__ import annotations

import torch

from .loss import LossOutput, create_loss_strategy
from .schedulers import CogVideoXDPMScheduler


class DPOSharedStep:
    def __init__(self, transformer, ref_transformer, beta: float = 1.0, scheduler=None, trainable=None, vae_encoder=None):
        """transformer / ref_transformer: policy and reference for the forward-only path (validation_step). trainable: a
        train_dit.LoRATrainableTransformer — the policy of training_step; its frozen base doubles as the reference model."""
        self.transformer, self.ref_transformer = transformer, ref_transformer
        self.trainable = trainable
        self.vae_encoder = vae_encoder            # I2V: encodes batch["image_emb"] into the first-frame condition
        self.scheduler = scheduler or CogVideoXDPMScheduler()
        self.loss_fn = create_loss_strategy(strategy="dpo", beta=beta)
        self.device = transformer.device

    def _prepare(self, batch: dict, generator=None, timesteps=None, noise=None):
        dev = self.device
        x_win = batch["x_win"].to(dev).permute(0, 2, 1, 3, 4).float()            # [B, F, C, H, W]
        x_lose = batch["x_lose"].to(dev).permute(0, 2, 1, 3, 4).float()
        prompt_emb = batch["prompt_emb"].to(dev)
        B = x_win.shape[0]
        if (self.transformer.config.patch_size_t or 1) > 1:
            # train/CogVideoX1.5-5B/03_train.py:134-144: the temporal-patch model needs even F, H, W; odd sizes are trimmed
            _, Fr, _, H, W = x_win.shape
            nF, nH, nW = Fr - Fr % 2, H - H % 2, W - W % 2
            if (nF, nH, nW) != (Fr, H, W):
                x_win = x_win[:, :nF, :, :nH, :nW].contiguous()
                x_lose = x_lose[:, :nF, :, :nH, :nW].contiguous()
                if noise is not None and tuple(noise.shape) != tuple(x_win.shape):
                    noise = noise[:, :nF, :, :nH, :nW].contiguous()
        if timesteps is None:
            timesteps = torch.randint(0, self.scheduler.num_train_timesteps, (B,), device=dev, generator=generator)
        if noise is None:
            noise = torch.randn(x_win.shape, device=dev, generator=generator)
        x_win_noisy = self.scheduler.add_noise(x_win, noise, timesteps)
        x_lose_noisy = self.scheduler.add_noise(x_lose, noise, timesteps)
        if self.transformer.config.in_channels == 2 * x_win.shape[2]:            # CogVideoX-5B-I2V: channel-concat the image condition
            img_cond = self._image_condition(batch, x_win, generator)
            x_win_noisy = torch.cat([x_win_noisy, img_cond], dim=2)
            x_lose_noisy = torch.cat([x_lose_noisy, img_cond], dim=2)
        pair = torch.cat([x_win_noisy, x_lose_noisy], dim=0)                     # one batch of 2B per model
        emb2 = torch.cat([prompt_emb, prompt_emb], dim=0)
        t2 = torch.cat([timesteps, timesteps], dim=0)
        v_win_target = self.scheduler.get_velocity(x_win, noise, timesteps)
        v_lose_target = self.scheduler.get_velocity(x_lose, noise, timesteps)
        return B, pair, emb2, t2, v_win_target.contiguous(), v_lose_target.contiguous()

    @torch.no_grad()
    def _image_condition(self, batch: dict, x_win: torch.Tensor, generator=None) -> torch.Tensor:
        """`img_cond` of train/CogVideoX-I2V-5B/03_train.py:119-130: the conditioning image resized to 8x the latent grid,
        VAE-encoded, sampled, times scaling_factor, as the first latent frame followed by zero frames; zeros without an image."""
        image_emb = batch.get("image_emb")
        if image_emb is None:
            return torch.zeros_like(x_win)
        if self.vae_encoder is None:
            raise RuntimeError("batch['image_emb'] needs vae_encoder=AutoencoderKLCogVideoXEncoder(...) (videogpa_b200.vae)")
        B, Fr, C, h, w = x_win.shape
        img = torch.nn.functional.interpolate(image_emb.to(self.device).float(), size=(h * 8, w * 8))
        dist = self.vae_encoder.encode(img.unsqueeze(2).to(self.vae_encoder.dtype)).latent_dist
        lat = dist.sample(generator=generator) * self.vae_encoder.config.scaling_factor      # [B, C, 1, h, w]
        cond = torch.zeros_like(x_win)
        cond[:, :1] = lat.permute(0, 2, 1, 3, 4).to(cond.dtype)
        return cond

    @torch.no_grad()
    def _shared_step(self, batch: dict, generator=None, timesteps=None, noise=None) -> LossOutput:
        B, pair, emb2, t2, v_win_target, v_lose_target = self._prepare(batch, generator, timesteps, noise)
        if self.trainable is not None:                  # the policy being trained (base + unmerged LoRA), without a graph
            v_pred = self.trainable(pair, emb2, t2)
        else:
            v_pred = self.transformer(pair, encoder_hidden_states=emb2, timestep=t2, return_dict=True).sample
        ref = self.ref_transformer if self.ref_transformer is not None else self.trainable.base
        v_ref = ref(pair, encoder_hidden_states=emb2, timestep=t2, return_dict=True).sample
        return self.loss_fn(v_pred[:B].contiguous(), v_pred[B:].contiguous(), v_ref[:B].contiguous(), v_ref[B:].contiguous(),
                            v_win_target, v_lose_target)

    def validation_step(self, batch: dict, batch_idx: int = 0, **kw) -> dict:
        """-> the scalars the reference logs (:189-201): val/loss, val/reward_margin, val/reward_accuracy."""
        out = self._shared_step(batch, **kw)
        return {"val/loss": out.loss, "val/reward_margin": out.reward_margin,
                "val/reward_accuracy": (out.reward_margin > 0).float().mean(), "loss_output": out}

    def training_step(self, batch: dict, batch_idx: int = 0, generator=None, timesteps=None, noise=None) -> torch.Tensor:
        """`training_step` of the reference (:159-187): returns the differentiable DPO loss; `loss.backward()` runs the hand-
        written backward kernels down to the LoRA factors (train_dit). The reference forwards run without a graph on the
        frozen base shared with the policy."""
        if self.trainable is None:
            raise RuntimeError("training_step needs `trainable=LoRATrainableTransformer(transformer)` (videogpa_b200.train_dit)")
        B, pair, emb2, t2, v_win_target, v_lose_target = self._prepare(batch, generator, timesteps, noise)
        with torch.no_grad():
            ref = self.ref_transformer if self.ref_transformer is not None else self.trainable.base
            v_ref = ref(pair, encoder_hidden_states=emb2, timestep=t2, return_dict=True).sample
        v_pred = self.trainable(pair, emb2, t2)
        out = self.loss_fn(v_pred[:B].contiguous(), v_pred[B:].contiguous(), v_ref[:B].contiguous(), v_ref[B:].contiguous(),
                           v_win_target, v_lose_target)
        self.last_output = out
        return out.loss

    def fit_step(self, batch: dict, optimizer, process_group=None, **kw) -> float:
        """zero_grad -> training_step -> backward -> (DDP gradient average) -> optimizer.step: what the Lightning trainer
        (`strategy="ddp"`, 03_train.py:258-266) does around training_step. With torch.distributed initialised every rank
        runs its own pairs and the LoRA gradients are all-reduced (parallel.average_gradients)."""
        import torch.distributed as dist
        optimizer.zero_grad(set_to_none=True)
        loss = self.training_step(batch, **kw)
        loss.backward()
        if dist.is_available() and dist.is_initialized():
            from .parallel import average_gradients
            average_gradients(self.trainable.parameters(), group=process_group)
        optimizer.step()
        return float(loss.detach())

    def configure_optimizers(self, lr: float = 5e-6):
        """`torch.optim.AdamW(self.transformer.parameters(), lr=5e-6)` over the trainable (LoRA) parameters (:207-210)."""
        return torch.optim.AdamW(self.trainable.parameters(), lr=lr)
