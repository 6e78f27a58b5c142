"""Latent encoding step of the reference's `train/*/02_encode.py` on the sm_100a VAE encoder — SURVEY.md §8 row f-4.

Mirrors the tensor half of `encode_video_latent` (`train/CogVideoX-5B/02_encode.py:97-123`,
`train/CogVideoX1.5-5B/02_encode.py:100-121`): sample a fixed number of frames, scale uint8 frames to [0, 1] (the reference
does NOT map them to [-1, 1]), `vae.encode(video).latent_dist.sample()`, move the latent to the host. The 5B scripts store
the latent unscaled, the 1.5 script multiplies by `vae.config.scaling_factor` — both kept, chosen by `scale_latents`.
`encode_text_condition` mirrors `encode_text_condition` of the same script (:69-93) on videogpa_b200.t5.T5EncoderModel.
Video decoding (decord) is outside this build: callers pass the decoded frames.
"""
from __future__ import annotations

import numpy as np
import torch


def select_frame_indices(total_frames: int, num_frames: int = 49) -> np.ndarray:
    """`load_video_frames_tensor` (train/CogVideoX-5B/02_encode.py:55-60): every frame when the clip is shorter than
    `num_frames`, else `np.linspace(0, total - 1, num_frames).astype(int)` (truncation, duplicates allowed)."""
    if total_frames < num_frames:
        return np.arange(0, total_frames).astype(int)
    return np.linspace(0, total_frames - 1, num_frames).astype(int)


def frames_to_video_tensor(frames, num_frames: int = 49, device="cuda") -> torch.Tensor:
    """frames [F, H, W, 3] uint8 (numpy or tensor) -> [3, F', H, W] float in [0, 1] on `device` (02_encode.py:61-63)."""
    fr = torch.as_tensor(frames)
    if fr.dim() != 4 or fr.shape[-1] != 3:
        raise RuntimeError(f"frames must be [F, H, W, 3], got {tuple(fr.shape)}")
    idx = torch.from_numpy(select_frame_indices(fr.shape[0], num_frames))
    return (fr[idx].float() / 255.0).permute(3, 0, 1, 2).to(device)


@torch.no_grad()
def encode_video_latent(vae_encoder, frames, num_frames: int = 49, scale_latents: bool = False,
                        generator: torch.Generator | None = None) -> torch.Tensor:
    """-> latent [C, T', h, w] on the host, as the reference saves it (`latent_dist.sample().squeeze(0).cpu()`)."""
    video = frames_to_video_tensor(frames, num_frames, vae_encoder.device).unsqueeze(0).to(vae_encoder.dtype)
    latent = vae_encoder.encode(video).latent_dist.sample(generator=generator)
    if scale_latents:
        latent = latent * vae_encoder.config.scaling_factor
    return latent.squeeze(0).cpu()


@torch.no_grad()
def encode_text_condition(text_encoder, input_ids: torch.Tensor) -> dict:
    """`encode_text_condition` (train/CogVideoX-5B/02_encode.py:69-93) after tokenisation: ids [1, 226] (max-length padded,
    truncated) -> {"encoder_hidden_states": [226, d_model] on the host}, the dict the script saves as `cond_<group>.pt`."""
    if input_ids.dim() != 2 or input_ids.shape[0] != 1:
        raise RuntimeError("input_ids must be [1, S]")
    emb = text_encoder(input_ids.to(text_encoder.device))[0].squeeze(0).cpu()
    return {"encoder_hidden_states": emb}
