"""`generate/CogVideoX1.5-5B.py` of the reference on the sm_100a kernels — same flags.

Reference: generate/CogVideoX1.5-5B.py:85,102-111: the CogVideoX-5B flags plus `--lora_weight (0.2)` (absolute LoRA
scaling, :32-35), `--height (768) --width (1360) --num_frames (81)`, fps 16 and `use_dynamic_cfg=True`. The 1.5
transformer uses temporal patching (patch_size_t = 2): videogpa_b200.transformer handles it and the pipeline pads /
drops the extra latent frame like diffusers. Everything else (prompt JSON forms, output layout, resume, per-prompt
error handling, `--synthetic`) is shared with videogpa_b200.generate.cogvideox_5b.
"""
from __future__ import annotations

import os
from pathlib import Path

import torch

from . import cogvideox_5b as base


def build_parser():
    p = base.build_parser()
    p.description = "CogVideoX1.5-5B T2V generation"
    p.set_defaults(base_model="THUDM/CogVideoX1.5-5B", fps=16, height=768, width=1360, num_frames=81)
    p.add_argument("--lora_weight", type=float, default=0.2)
    return p


def generate(args):
    device = torch.device(f"cuda:{args.gpu_id}")
    torch.cuda.set_device(device)
    print(f"Loading base model: {args.base_model}")
    if args.synthetic:
        from ..pipeline import CogVideoXDenoisePipeline
        from ..schedulers import CogVideoXDPMScheduler
        from ..transformer import CogVideoXTransformer3D, TransformerConfig
        from ..vae import AutoencoderKLCogVideoXDecoder, VAEDecoderConfig
        cfg = TransformerConfig.cogvideox1_5_5b()
        cfg.num_layers = args.synthetic
        vae = AutoencoderKLCogVideoXDecoder.random_init(VAEDecoderConfig(sample_height=args.height, sample_width=args.width), seed=5, device=device)
        vae.enable_tiling(); vae.enable_slicing()
        pipe = CogVideoXDenoisePipeline(CogVideoXTransformer3D.random_init(cfg, seed=1234, device=device), CogVideoXDPMScheduler(), vae=vae,
                                        vae_scaling_factor=vae.config.scaling_factor)
        prompts = base._SyntheticPrompts(cfg.text_embed_dim, device)
    else:
        pipe, prompts = base.build_pipeline(args, device)
    if args.lora_path:
        if not os.path.exists(args.lora_path):
            print(f"LoRA path not found: {args.lora_path}, using base model")
        else:
            from ..lora import merge_lora
            print(f"Mounting LoRA: {args.lora_path} (scaling {args.lora_weight})")
            merge_lora(pipe.transformer, args.lora_path, scaling=args.lora_weight)          # absolute scaling, :32-35
            print("LoRA merged.")
    tasks = base.load_tasks(args.prompt_json, args.num_prompts)
    if tasks is None:
        print("Unsupported JSON format")
        return
    print(f"Generating {len(tasks)} prompts, seed={args.seed}")
    output_root = Path(args.output_dir)
    output_root.mkdir(parents=True, exist_ok=True)
    # The reference calls pipe(...) without max_sequence_length, so diffusers pads prompts to its default of 226 tokens
    # (train/CogVideoX1.5-5B/02_encode.py uses max_length=226 too). T5 runs unmasked, so the pad count changes every token:
    # config.max_text_seq_length (224 for 1.5) only sizes a learned positional embedding, which 1.5 does not have.
    max_len = 226
    negative = prompts("", max_len)
    for idx, item in enumerate(tasks):
        text_prompt = item.get("text_prompt", item.get("prompt", "")).strip()
        if not text_prompt:
            continue
        group_id, video_path = base.video_path_for(output_root, item, idx, args.seed)
        video_path.parent.mkdir(parents=True, exist_ok=True)
        if video_path.exists():
            print(f"[{idx+1}/{len(tasks)}] Skip existing: {group_id}")
            continue
        print(f"[{idx+1}/{len(tasks)}] Generating: {group_id}")
        try:
            generator = torch.Generator(device=device).manual_seed(args.seed)
            frames = pipe(prompts(text_prompt, max_len), negative, num_frames=args.num_frames, height=args.height, width=args.width,
                          num_inference_steps=args.num_inference_steps, guidance_scale=args.guidance_scale, use_dynamic_cfg=True,
                          generator=generator, output_type="pt")
            base.export_to_video(frames[0], str(video_path), fps=args.fps)
        except Exception as e:                      # noqa: BLE001
            print(f"  Failed: {e}")
        torch.cuda.empty_cache()
    print("Done.")


def main(argv=None):
    generate(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
