"""`train/CogVideoX-5B/02_encode.py` (and the 1.5 variant) of the reference: turn the scored metadata into training inputs — one
text-condition file per group and one VAE latent per video — and write the metadata JSON `DPODataset` reads.

Reference: train/CogVideoX-5B/02_encode.py:69-93 (`cond_<group_id>.pt` = {"encoder_hidden_states": [226, 4096]}: tokenizer with
`padding="max_length", max_length=226, truncation=True`, `text_encoder(ids)[0].squeeze(0).cpu()`), :95-123 (`latent_<group_id>_<stem>.pt`:
49 frames of `<BASE_PATH>/<video_path>` in [0, 1], `vae.encode(...).latent_dist.sample().squeeze(0).cpu()`, unscaled),
train/CogVideoX1.5-5B/02_encode.py:100-121 (the same with `* vae.config.scaling_factor`), :128-214 (worker: groups without prompt or
videos skipped, paths stored relative to BASE_PATH, groups without any encoded video dropped, output group = {group_id, text_prompt,
videos}), :219-262 (input = a list or {"t2v_groups" | "groups": [...]}, round-robin split over the GPUs, output {"groups": [...]}).

The encoders are this package's (`vae.AutoencoderKLCogVideoXEncoder`, `t5.T5EncoderModel`) or anything with the same call shape;
the tokenizer is host string work (`tokenize_fn(prompt) -> ids [1, 226]`, e.g. transformers' AutoTokenizer as in
generate.cogvideox_5b._T5Prompts); videos are decoded by `video_io.load_video_frames_tensor`. One process per GPU under torchrun.
"""
from __future__ import annotations

import logging
import os
from pathlib import Path

import torch

from ..encode import encode_text_condition
from .preference_pair import safe_load_json, safe_save_json


def extract_groups(data):
    """02_encode.py:232-237."""
    if isinstance(data, list):
        return data
    if isinstance(data, dict):
        return data.get("t2v_groups", []) or data.get("groups", [])
    return []


@torch.no_grad()
def encode_video_file(vae_encoder, video_full_path, num_frames: int = 49, scale_latents: bool = False, generator=None) -> torch.Tensor:
    """02_encode.py:95-115 -> latent [C, T', h, w] on the host."""
    from ..video_io import load_video_frames_tensor
    video = load_video_frames_tensor(str(video_full_path), num_frames, vae_encoder.device).unsqueeze(0).to(vae_encoder.dtype)
    latent = vae_encoder.encode(video).latent_dist.sample(generator=generator)
    if scale_latents:                                            # train/CogVideoX1.5-5B/02_encode.py:113-114
        latent = latent * vae_encoder.config.scaling_factor
    return latent.squeeze(0).cpu()


def load_input_image_tensor(image_path) -> torch.Tensor | None:
    """train/CogVideoX-I2V-5B/02_encode.py:65-69: RGB image -> float [3, H, W] in [0, 1] (None when the file is missing)."""
    import numpy as np
    from PIL import Image
    p = Path(image_path)
    if not p.exists():
        return None
    arr = np.array(Image.open(p).convert("RGB")).astype(np.float32) / 255.0
    return torch.from_numpy(arr).permute(2, 0, 1)


def encode_groups(groups_chunk: list, base_path, latent_root, vae_encoder, text_encoder, tokenize_fn, num_frames: int = 49,
                  scale_latents: bool = False, generator=None, tag: str = "Worker-0", image_condition: bool = False, sub_folder: str = "") -> list:
    """The loop of `gpu_worker` (02_encode.py:161-214). image_condition: the I2V variant (train/CogVideoX-I2V-5B/02_encode.py:72-99,
    :144-176) — groups need an `image_path`, the conditioning image is stored as `image_embeds` in the condition file, files go under
    `<latent_root>/<sub_folder>` ("processed" there) and the whole input group is kept, not only {group_id, text_prompt, videos}."""
    base_path, out = Path(base_path), Path(latent_root) / sub_folder
    out.mkdir(parents=True, exist_ok=True)
    processed = []
    for group in groups_chunk:
        prompt, group_id, entries = group.get("text_prompt"), group.get("group_id"), group.get("videos", [])
        if not prompt or not entries or (image_condition and not group.get("image_path")):
            logging.warning(f"Group {group_id} missing prompt, videos or image, skipped.")
            continue
        try:
            cond_path = out / f"cond_{group_id}.pt"
            cond = encode_text_condition(text_encoder, tokenize_fn(prompt))
            if image_condition:
                img = load_input_image_tensor(base_path / group["image_path"])
                if img is not None:
                    cond["image_embeds"] = img
            torch.save(cond, cond_path)
            videos = []
            for entry in entries:
                rel = entry.get("video_path")
                if not rel:
                    continue
                try:
                    full = base_path / rel
                    if not full.exists():
                        raise FileNotFoundError(f"Video not found: {full}")
                    latent = encode_video_file(vae_encoder, full, num_frames, scale_latents, generator)
                    latent_path = out / f"latent_{group_id}_{Path(rel).stem}.pt"
                    torch.save(latent, latent_path)
                    e = entry.copy()
                    e["condition_path"] = str(cond_path.relative_to(base_path))
                    e["latent_path"] = str(latent_path.relative_to(base_path))
                    videos.append(e)
                except Exception as ex:                            # noqa: BLE001
                    logging.error(f"Video {rel} Encoding Error: {ex}")
                    continue
            if videos:
                if image_condition:
                    kept = dict(group)
                    kept["videos"] = videos
                    processed.append(kept)
                else:
                    processed.append({"group_id": group_id, "text_prompt": prompt, "videos": videos})
        except Exception as ex:                                    # noqa: BLE001
            logging.error(f"Group {group_id} Encoding Error: {ex}")
            continue
    return processed


def process_t2v_encoding(input_json: str, output_json: str, base_path: str, vae_encoder, text_encoder, tokenize_fn,
                         latent_root: str | None = None, num_frames: int = 49, scale_latents: bool = False, generator=None,
                         image_condition: bool = False, sub_folder: str = ""):
    """02_encode.py:219-262. -> {"groups": [...]} as written to `output_json` (rank 0), None otherwise."""
    data = safe_load_json(input_json)
    if not data:
        logging.error(f"Input JSON not found or invalid: {input_json}")
        return None
    all_groups = extract_groups(data)
    if not all_groups:
        logging.warning("No valid T2V groups found in input JSON.")
        return None
    latent_root = latent_root or os.path.join(base_path, "t2v_latent")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = 0
    if world > 1:
        from ..parallel import init_from_env
        rank, world, _ = init_from_env()
    mine = encode_groups(all_groups[rank::world], base_path, latent_root, vae_encoder, text_encoder, tokenize_fn, num_frames, scale_latents,
                         generator, tag=f"Worker-{rank}", image_condition=image_condition, sub_folder=sub_folder)
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        mine = [g for sub in gathered for g in sub]
    if rank != 0:
        return None
    if not mine:
        logging.warning("No groups processed successfully.")
        return None
    payload = {"groups": mine}
    safe_save_json(output_json, payload)
    logging.info(f"T2V Encoding Complete, processed {len(mine)} groups, results saved to {output_json}")
    return payload
