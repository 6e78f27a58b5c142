"""`replicate.py` of the reference (BASELINE.json configs[2]: CogVideoX-5B-I2V + VideoGPA LoRA, prompts sharded over the GPUs of a box)
on the sm_100a pipeline.

Reference: replicate.py:11-43 (the `RUN_*` / `PROMPT_JSON` / `DL3DV_BASE_DIR` environment configuration, same names and defaults),
:46-97 (`extract_pure_hash_from_json_key`: the middle part of `1K/<hash>/images_8`; `find_dl3dv_first_frame`: `<base>/<key>/frame_00001.png`),
:100-146 (first `RUN_NUM_PROMPTS` entries of the caption JSON, round-robin shard `items[i::num_gpus]`, one worker per GPU),
:149-250 (worker: I2V pipeline with VAE tiling + slicing; modes dpo / dpo_epipolar / sft attach the LoRA adapter UNMERGED and, per work
item, set `module.scaling = lora_weight * lora_alpha / r`; other modes run the base model with weight 0.0; first frame resized to
1080x720; `seed_<seed>_<mode>_w<weight>.mp4` / `seed_<seed>_original.mp4` under `<output_dir>/<hash>/`; existing files skipped; a failing
item is printed and skipped).

One process per GPU under torchrun (rank r takes `items[r::world]`) instead of `mp.Process`; the adapter strength goes through
`lora.attach_lora(...).set_weight(w)` (fused weights rebuilt from the pristine base, one rounding); nothing is exchanged between ranks:
the prompt shard has no data-path collective. `plan_jobs` is the pure bookkeeping (which files a worker will write), shared with the tests.
"""
from __future__ import annotations

import argparse
import json
import os
from pathlib import Path

LORA_MODES = ("dpo", "dpo_epipolar", "sft")


def parse_int_list_env(name, default):
    raw = os.getenv(name)
    if not raw:
        return list(default)
    return [int(item.strip()) for item in raw.split(",") if item.strip()]


def build_config(here: str | None = None) -> dict:
    """replicate.py:19-43."""
    here = here or os.getcwd()
    return {
        "devices": parse_int_list_env("RUN_DEVICES", [0]),
        "mode": os.getenv("RUN_MODE", "dpo"),
        "weight_list": [1.0],
        "base_model": "THUDM/CogVideoX-5B-I2V",
        "lora_path": os.getenv("RUN_LORA_PATH", os.path.join(here, "checkpoints/VideoGPA-I2V-lora")),
        "prompt_json": os.getenv("PROMPT_JSON", os.path.join(here, "dl3dv_video_captions/captions_1K.json")),
        "dl3dv_base_dir": os.getenv("DL3DV_BASE_DIR", "/datasets/DL3DV-10K"),
        "output_dir": os.getenv("RUN_OUTPUT_DIR", os.path.join(here, "output/replicate")),
        "num_prompts": int(os.getenv("RUN_NUM_PROMPTS", "100")),
        "seeds_per_prompt": parse_int_list_env("RUN_SEEDS", [456]),
        "num_inference_steps": 50,
        "guidance_scale": 6.0,
        "fps": 8,
    }


def extract_pure_hash_from_json_key(json_key: str) -> str:
    """replicate.py:46-63: `1K/<hash>/images_8` -> `<hash>`; any other key with its separators replaced by `_`."""
    try:
        json_key = json_key.strip()
        parts = json_key.split("/")
        pure_hash = parts[1] if len(parts) == 3 else json_key.replace("/", "_").replace("\\", "_")
        if not pure_hash:
            raise ValueError("Extracted hash string is empty, cannot use as storage folder")
        return pure_hash
    except Exception as e:                                     # noqa: BLE001
        raise RuntimeError(f"Failed to extract pure hash string: {str(e)}") from e


def find_dl3dv_first_frame(json_key: str, dl3dv_base_dir):
    """replicate.py:66-97 -> (first frame path, pure hash, frame folder)."""
    json_key = json_key.strip()
    if not json_key:
        raise ValueError("JSON Key is empty, cannot construct path")
    target = Path(dl3dv_base_dir) / json_key
    if not target.exists() or not target.is_dir():
        raise FileNotFoundError(f"Frame folder not found under DL3DV directory: {target}")
    first = target / "frame_00001.png"
    if not first.exists():
        raise FileNotFoundError("Standard first frame not found in frame folder: frame_00001.png")
    return first, extract_pure_hash_from_json_key(json_key), target


def select_items(config: dict) -> list:
    """replicate.py:108-117: the first `num_prompts` (key, caption) entries of the caption JSON."""
    with open(config["prompt_json"], "r", encoding="utf-8") as f:
        prompt_dict = json.load(f)
    return list(prompt_dict.items())[: config["num_prompts"]]


def weights_for(config: dict) -> list:
    return list(config["weight_list"]) if config["mode"] in LORA_MODES else [0.0]


def video_filename(mode: str, seed: int, lora_weight: float) -> str:
    """replicate.py:218-221."""
    return f"seed_{seed}_{mode}_w{lora_weight}.mp4" if mode in LORA_MODES else f"seed_{seed}_original.mp4"


def plan_jobs(json_items: list, config: dict, log=print) -> list:
    """The files one worker will write, in the reference's loop order (item -> weight -> seed); invalid items are skipped with
    the reference's messages. -> [dict(json_key, prompt, first_frame, pure_hash, lora_weight, seed, path)]."""
    jobs = []
    for json_key, text_prompt in json_items:
        text_prompt = text_prompt.strip()
        if not text_prompt:
            log(f"skipping invalid entry (empty prompt): {json_key}")
            continue
        try:
            first_frame, pure_hash, _ = find_dl3dv_first_frame(json_key, config["dl3dv_base_dir"])
        except Exception as e:                                 # noqa: BLE001
            log(f"skipping entry: {e}")
            continue
        out_dir = Path(config["output_dir"]) / pure_hash
        for w in weights_for(config):
            for seed in config["seeds_per_prompt"]:
                jobs.append(dict(json_key=json_key, prompt=text_prompt, first_frame=first_frame, pure_hash=pure_hash, lora_weight=w,
                                 seed=seed, path=out_dir / video_filename(config["mode"], seed, w)))
    return jobs


def worker(rank: int, gpu_id: int, json_items: list, config: dict, synthetic: int = 0) -> int:
    """replicate.py:149-250 on one GPU. -> number of videos written."""
    import torch
    from PIL import Image
    from .generate import cogvideox_5b as base
    from .generate import cogvideox_5b_i2v as i2v
    from .lora import attach_lora
    device = torch.device(f"cuda:{gpu_id}")
    torch.cuda.set_device(device)
    mode = config["mode"]
    print(f"GPU {gpu_id} | Starting generation | Mode: {mode} | Weights: {weights_for(config)}")
    args = i2v.build_parser().parse_args(["--prompt_json", config["prompt_json"], "--output_dir", config["output_dir"],
                                          "--base_model", config["base_model"], "--gpu_id", str(gpu_id)])
    args.synthetic = synthetic
    if synthetic:
        args.height, args.width, args.num_frames = 96, 160, 9                       # a grid a few-block random model handles in seconds
        from .pipeline import CogVideoXDenoisePipeline
        from .schedulers import CogVideoXDDIMScheduler
        from .transformer import CogVideoXTransformer3D, TransformerConfig
        from .vae import AutoencoderKLCogVideoXDecoder, AutoencoderKLCogVideoXEncoder, VAEDecoderConfig
        cfg = TransformerConfig.cogvideox_5b_i2v()
        cfg.num_layers = synthetic
        cfg.sample_height, cfg.sample_width, cfg.sample_frames = args.height // 8, args.width // 8, args.num_frames
        vae = AutoencoderKLCogVideoXDecoder.random_init(VAEDecoderConfig(), seed=5, device=device)
        pipe = CogVideoXDenoisePipeline(CogVideoXTransformer3D.random_init(cfg, seed=1234, device=device), CogVideoXDDIMScheduler(), vae=vae,
                                        vae_scaling_factor=vae.config.scaling_factor)
        pipe.vae_encoder = AutoencoderKLCogVideoXEncoder.random_init(VAEDecoderConfig(), seed=6, device=device)
        prompts = base._SyntheticPrompts(cfg.text_embed_dim, device)
    else:
        pipe, prompts = base.build_pipeline(args, device, with_encoder=True)
        pipe.scheduler = i2v.checkpoint_scheduler(args.base_model)    # CogVideoXImageToVideoPipeline keeps the checkpoint's scheduler
    pipe.vae.enable_tiling(); pipe.vae.enable_slicing()
    handle = None
    if mode in LORA_MODES and config["lora_path"]:
        print(f"GPU {gpu_id} loading LoRA weights: {config['lora_path']}")
        try:
            handle = attach_lora(pipe.transformer, config["lora_path"], weight=None)       # strength is set per work item below
        except Exception as e:                                 # noqa: BLE001
            print(f"GPU {gpu_id} LoRA loading failed: {e}")
            return 0
    negative = prompts("")
    written, current_weight = 0, None
    for job in plan_jobs(json_items, config, log=lambda m: print(f"GPU {gpu_id} {m}")):
        job["path"].parent.mkdir(parents=True, exist_ok=True)
        if handle is not None and job["lora_weight"] != current_weight:
            handle.set_weight(job["lora_weight"])                                           # :208-213
            current_weight = job["lora_weight"]
        if job["path"].exists():
            print(f"GPU {gpu_id} video already exists, skipping: {job['path'].name}")
            continue
        try:
            image = Image.open(str(job["first_frame"])).convert("RGB").resize((1080, 720))  # :199-200
            generator = torch.Generator(device=device).manual_seed(job["seed"])
            print(f"GPU {gpu_id} generating {job['pure_hash']} - Seed {job['seed']}...")
            _, F_, C, h, w = pipe.latent_shape(1, args.num_frames, args.height, args.width)
            img_lat = i2v.first_frame_latent(image, (F_, C, h, w), device, pipe.vae_encoder, pipe.vae.config.scaling_factor,
                                             generator=generator, height=args.height, width=args.width)
            frames = pipe(prompts(job["prompt"]), negative, num_frames=args.num_frames, height=args.height, width=args.width,
                          num_inference_steps=config["num_inference_steps"], guidance_scale=config["guidance_scale"], generator=generator,
                          image_latents=img_lat, output_type="pt")
            base.export_to_video(frames[0], str(job["path"]), fps=config["fps"])
            print(f"GPU {gpu_id} saved: {job['path'].name}")
            written += 1
        except Exception as e:                                 # noqa: BLE001
            print(f"GPU {gpu_id} generation failed {job['pure_hash']} - Seed {job['seed']}: {e}")
            continue
    torch.cuda.empty_cache()
    print(f"GPU {gpu_id} tasks completed")
    return written


def main(argv=None, config: dict | None = None) -> int:
    ap = argparse.ArgumentParser(description="VideoGPA replicate driver (I2V generation sharded over GPUs)")
    ap.add_argument("--synthetic", type=int, default=0, help="N > 0: N-block random-weight I2V model on a small grid (no checkpoint needed)")
    a = ap.parse_args(argv)
    config = config or build_config()
    Path(config["output_dir"]).mkdir(parents=True, exist_ok=True)
    if not Path(config["dl3dv_base_dir"]).exists():
        print(f"DL3DV root directory does not exist: {config['dl3dv_base_dir']}")
        return 0
    try:
        selected = select_items(config)
        print(f"Selected {len(selected)} entries for generation")
    except Exception as e:                                     # noqa: BLE001
        print(f"Failed to load JSON file: {e}")
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    gpu = config["devices"][local] if local < len(config["devices"]) else local
    n = worker(rank, gpu, selected[rank::world], config, synthetic=a.synthetic)               # :119-120 items[i::num_gpus]
    print(f"\nAll generation tasks for mode {config['mode']} completed on rank {rank}: {n} videos")
    print(f"Video output directory: {config['output_dir']}/<pure_hash>/seed_xxx.mp4")
    return n


if __name__ == "__main__":
    main()
