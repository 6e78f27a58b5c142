"""`generate/CogVideoX-5B-I2V.py` of the reference on the sm_100a kernels — same flags.

Reference: generate/CogVideoX-5B-I2V.py:12-115: the T2V flags plus `--base_dir` (root for relative image paths); tasks are
`(group_id, {text_prompt, image_prompt | image_path | input_image_path})`; items without a prompt or an image are
skipped, a missing image prints `Image not found ... skipping`; the checkpoint's default (DDIM) scheduler is kept.
The I2V transformer (in_channels 32, learned positional embedding) and the channel-concat of the encoded first frame run
on videogpa_b200 kernels. The first-frame latent is `vae.encode(image).latent_dist.sample(generator) * scaling_factor`
as in the library's I2V `prepare_latents` (resize to width x height, [-1, 1]), drawn from the prompt's generator BEFORE
the initial noise; the remaining F-1 latent frames are zero. `--synthetic N` runs the same flow with seeded random
transformer / VAE weights so the CLI contract can be exercised without checkpoints.
"""
from __future__ import annotations

import json
import os
from pathlib import Path

import torch

from . import cogvideox_5b as base


def load_tasks(prompt_json: str, num_prompts: int | None):
    """generate/CogVideoX-5B-I2V.py:36-48."""
    with open(prompt_json, "r", encoding="utf-8") as f:
        raw = json.load(f)
    if isinstance(raw, dict):
        tasks = list(raw.items())
    elif isinstance(raw, list):
        tasks = [(item.get("group_id", i), item) for i, item in enumerate(raw)]
    else:
        return None
    return tasks[:num_prompts] if num_prompts else tasks


def resolve_image(item: dict, base_dir: str | None) -> str:
    """generate/CogVideoX-5B-I2V.py:56-64."""
    image_path = item.get("image_prompt", item.get("image_path", item.get("input_image_path", "")))
    if image_path and not Path(image_path).exists() and base_dir:
        image_path = str(Path(base_dir) / image_path)
    return image_path


def load_image_tensor(image_path, height: int, width: int) -> torch.Tensor:
    """diffusers `load_image` + VideoProcessor.preprocess: RGB, resized to (width, height) (lanczos), [-1, 1], [3, H, W]."""
    import numpy as np
    from PIL import Image
    img = image_path if hasattr(image_path, "resize") else Image.open(image_path)          # a path or an already loaded PIL image
    img = img.convert("RGB").resize((width, height), Image.LANCZOS)
    return torch.from_numpy(np.asarray(img).copy()).permute(2, 0, 1).float() / 127.5 - 1.0


def first_frame_latent(image_path, shape, device, encoder, scaling_factor: float, generator=None,
                       height: int | None = None, width: int | None = None) -> torch.Tensor:
    """-> image latents [1, F, 16, h, w]: the encoded first frame followed by F-1 zero frames (App. A.4)."""
    F_, C, h, w = shape
    if encoder is None:
        raise RuntimeError("the I2V pipeline needs the VAE encoder")
    img = load_image_tensor(image_path, height or h * 8, width or w * 8).to(device=device, dtype=torch.bfloat16)
    first = encoder.encode(img[None, :, None]).latent_dist.sample(generator=generator)[0] * scaling_factor     # [C, 1, h, w]
    if tuple(first.shape) != (C, 1, h, w):
        raise RuntimeError(f"encoded first frame has shape {tuple(first.shape)}, expected {(C, 1, h, w)}")
    lat = torch.zeros(1, F_, C, h, w, dtype=torch.bfloat16, device=device)
    lat[0, 0] = first[:, 0].to(torch.bfloat16)
    return lat


def checkpoint_scheduler(base_model):
    """The scheduler `CogVideoXImageToVideoPipeline.from_pretrained` builds: class and settings from `<base_model>/scheduler/
    scheduler_config.json` (CogVideoX-5B-I2V ships CogVideoXDDIMScheduler)."""
    import json
    from ..schedulers import CogVideoXDDIMScheduler, CogVideoXDPMScheduler
    sc_file = Path(base_model) / "scheduler" / "scheduler_config.json"
    if not sc_file.is_file():
        raise RuntimeError(f"{sc_file} not found")
    sc = json.loads(sc_file.read_text())
    known = {"CogVideoXDDIMScheduler": CogVideoXDDIMScheduler, "CogVideoXDPMScheduler": CogVideoXDPMScheduler}
    name = sc.get("_class_name", "CogVideoXDDIMScheduler")
    if name not in known:
        raise RuntimeError(f"scheduler class {name!r} is not implemented")
    return known[name].from_config(sc)


def build_parser():
    p = base.build_parser()
    p.description = "CogVideoX-5B I2V generation"
    p.set_defaults(base_model="THUDM/CogVideoX-5B-I2V")
    p.add_argument("--base_dir", type=str, default=None, help="Base dir for relative image paths")
    return p


def generate(args):
    device = torch.device(f"cuda:{args.gpu_id}")
    torch.cuda.set_device(device)
    print(f"Loading base model: {args.base_model}")
    from ..schedulers import CogVideoXDDIMScheduler
    if args.synthetic:
        from ..pipeline import CogVideoXDenoisePipeline
        from ..transformer import CogVideoXTransformer3D, TransformerConfig
        from ..vae import AutoencoderKLCogVideoXDecoder, AutoencoderKLCogVideoXEncoder, VAEDecoderConfig
        cfg = TransformerConfig.cogvideox_5b_i2v()
        cfg.num_layers = args.synthetic
        cfg.sample_height, cfg.sample_width = args.height // 8, args.width // 8
        cfg.sample_frames = args.num_frames
        vae = AutoencoderKLCogVideoXDecoder.random_init(VAEDecoderConfig(), seed=5, device=device)
        vae.enable_tiling(); vae.enable_slicing()
        pipe = CogVideoXDenoisePipeline(CogVideoXTransformer3D.random_init(cfg, seed=1234, device=device), CogVideoXDDIMScheduler(), vae=vae,
                                        vae_scaling_factor=vae.config.scaling_factor)
        pipe.vae_encoder = AutoencoderKLCogVideoXEncoder.random_init(VAEDecoderConfig(), seed=6, device=device)
        pipe.vae_encoder.enable_tiling(); pipe.vae_encoder.enable_slicing()
        prompts = base._SyntheticPrompts(cfg.text_embed_dim, device)
    else:
        pipe, prompts = base.build_pipeline(args, device, with_encoder=True)
        pipe.scheduler = checkpoint_scheduler(args.base_model)      # the I2V script keeps the checkpoint's own scheduler (:16-19, no swap)
    if args.lora_path:
        if not os.path.exists(args.lora_path):
            print(f"LoRA path not found: {args.lora_path}, using base model")
        else:
            from ..lora import merge_lora
            print(f"Mounting LoRA: {args.lora_path}")
            merge_lora(pipe.transformer, args.lora_path)
            print("LoRA merged.")
    tasks = load_tasks(args.prompt_json, args.num_prompts)
    if tasks is None:
        print("Unsupported JSON format")
        return
    print(f"Generating {len(tasks)} prompts, seed={args.seed}")
    output_root = Path(args.output_dir)
    output_root.mkdir(parents=True, exist_ok=True)
    negative = prompts("")
    for idx, (group_id, item) in enumerate(tasks):
        group_id = str(group_id).replace("/", "_")
        text_prompt = item.get("text_prompt", item.get("prompt", "")).strip()
        image_path = resolve_image(item, args.base_dir)
        if not text_prompt or not image_path:
            continue
        if not Path(image_path).exists():
            print(f"[{idx+1}/{len(tasks)}] Image not found: {image_path}, skipping")
            continue
        video_path = output_root / group_id / f"seed_{args.seed}.mp4"
        video_path.parent.mkdir(parents=True, exist_ok=True)
        if video_path.exists():
            print(f"[{idx+1}/{len(tasks)}] Skip existing: {group_id}")
            continue
        print(f"[{idx+1}/{len(tasks)}] Generating: {group_id}")
        try:
            generator = torch.Generator(device=device).manual_seed(args.seed)
            _, F_, C, h, w = pipe.latent_shape(1, args.num_frames, args.height, args.width)
            img_lat = first_frame_latent(image_path, (F_, C, h, w), device, pipe.vae_encoder, pipe.vae.config.scaling_factor,
                                         generator=generator, height=args.height, width=args.width)
            frames = pipe(prompts(text_prompt), negative, num_frames=args.num_frames, height=args.height, width=args.width,
                          num_inference_steps=args.num_inference_steps, guidance_scale=args.guidance_scale, generator=generator,
                          image_latents=img_lat, output_type="pt")
            base.export_to_video(frames[0], str(video_path), fps=args.fps)
        except Exception as e:                      # noqa: BLE001
            print(f"  Failed: {e}")
        torch.cuda.empty_cache()
    print("Done.")


def main(argv=None):
    generate(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
