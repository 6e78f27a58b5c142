"""`generate/CogVideoX-5B.py` of the reference on the sm_100a kernels — same flags, same output layout.

Reference: generate/CogVideoX-5B.py:11-99. Kept verbatim: `--base_model --prompt_json(required) --output_dir(required)
--lora_path --gpu_id --seed(42) --num_prompts --num_inference_steps(50) --guidance_scale(6.0) --fps(8)`; prompt JSON as
a dict `{group_id: "prompt" | {"text_prompt"/"prompt": ...}}` or a list of `{group_id, text_prompt}`; output
`<output_dir>/<group_id>/seed_<seed>.mp4`; an existing file is skipped (resume); a failing prompt prints `Failed: ...`
and the loop continues; a missing `--lora_path` prints a warning and runs the base model; the DPM scheduler with
trailing spacing replaces the checkpoint default; VAE tiling + slicing are enabled.

`--base_model` must be a local diffusers-layout directory (`transformer/`, `vae/`, `text_encoder/`, `tokenizer/`):
there is no network. The DiT and the VAE decoder run on videogpa_b200 kernels; the T5 text encoder is the
third-party `transformers` model (out of scope). Extra flag (not in the reference):
`--synthetic N` runs N-block random-init weights and hash-seeded prompt embeddings so the CLI contract can be
exercised without checkpoints.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
from pathlib import Path

import torch


def load_tasks(prompt_json: str, num_prompts: int | None):
    """generate/CogVideoX-5B.py:36-48."""
    with open(prompt_json, "r", encoding="utf-8") as f:
        raw = json.load(f)
    if isinstance(raw, dict):
        tasks = [{"group_id": k, "text_prompt": v if isinstance(v, str) else v.get("text_prompt", v.get("prompt", ""))}
                 for k, v in raw.items()]
    elif isinstance(raw, list):
        tasks = raw
    else:
        return None
    if num_prompts:
        tasks = tasks[:num_prompts]
    return tasks


def video_path_for(output_root: Path, item: dict, idx: int, seed: int) -> tuple[str, Path]:
    """generate/CogVideoX-5B.py:55-62."""
    group_id = str(item.get("group_id", idx)).replace("/", "_")
    return group_id, output_root / group_id / f"seed_{seed}.mp4"


def export_to_video(frames: torch.Tensor, path: str, fps: int) -> None:
    """frames [3, T, H, W] in [-1, 1] -> mp4 (diffusers.utils.export_to_video stand-in, OpenCV mp4v)."""
    import cv2
    import numpy as np
    vid = ((frames.float().clamp(-1, 1) + 1) * 127.5).round().to(torch.uint8).permute(1, 2, 3, 0).cpu().numpy()
    T, H, W, _ = vid.shape
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (W, H))
    if not wr.isOpened():
        raise RuntimeError(f"cannot open {path} for writing")
    for t in range(T):
        wr.write(np.ascontiguousarray(vid[t][:, :, ::-1]))
    wr.release()


def _load_safetensors_dir(d: Path) -> dict:
    from safetensors.torch import load_file
    files = sorted(d.glob("*.safetensors"))
    if not files:
        raise RuntimeError(f"no .safetensors weights under {d}")
    sd = {}
    for f in files:
        sd.update(load_file(str(f)))
    return sd


class _T5Prompts:
    """`encode_prompt` of the diffusers pipeline: tokenizer (transformers, host string work) + the T5 encoder on the
    sm_100a kernels (videogpa_b200.t5), 226 max-length-padded tokens, no attention mask."""

    def __init__(self, base: Path, device):
        from transformers import AutoTokenizer
        from ..t5 import T5Config, T5EncoderModel
        self.tok = AutoTokenizer.from_pretrained(str(base / "tokenizer"))
        tcfg = json.loads((base / "text_encoder" / "config.json").read_text())
        if tcfg.get("feed_forward_proj", "gated-gelu") != "gated-gelu":
            raise RuntimeError("only the gated-gelu T5 v1.1 text encoder is supported")
        known = T5Config.__dataclass_fields__.keys()
        self.enc = T5EncoderModel(T5Config(**{k: v for k, v in tcfg.items() if k in known}),
                                  _load_safetensors_dir(base / "text_encoder"), device=device)
        self.device = device

    @torch.no_grad()
    def __call__(self, prompt: str, max_len: int = 226) -> torch.Tensor:
        ids = self.tok(prompt, padding="max_length", max_length=max_len, truncation=True, add_special_tokens=True,
                       return_tensors="pt").input_ids.to(self.device)
        return self.enc(ids)[0].to(torch.bfloat16)


class _SyntheticPrompts:
    """Deterministic stand-in for the text encoder: embeddings seeded by the prompt's hash."""

    def __init__(self, dim: int, device):
        self.dim, self.device = dim, device

    def __call__(self, prompt: str, max_len: int = 226) -> torch.Tensor:
        seed = int.from_bytes(hashlib.sha256(prompt.encode()).digest()[:4], "little")
        g = torch.Generator().manual_seed(seed)
        return torch.randn(1, max_len, self.dim, generator=g).to(self.device, torch.bfloat16)


def build_pipeline(args, device, with_encoder: bool = False):
    """-> (pipeline, prompt encoder). `with_encoder` also builds the VAE encoder (I2V first frame) as `pipe.vae_encoder`."""
    from ..pipeline import CogVideoXDenoisePipeline
    from ..schedulers import CogVideoXDPMScheduler
    from ..transformer import CogVideoXTransformer3D, TransformerConfig
    from ..vae import AutoencoderKLCogVideoXDecoder, VAEDecoderConfig
    if args.synthetic:
        cfg = TransformerConfig.cogvideox_5b()
        cfg.num_layers = args.synthetic
        transformer = CogVideoXTransformer3D.random_init(cfg, seed=1234, device=device)
        vae = AutoencoderKLCogVideoXDecoder.random_init(VAEDecoderConfig(), seed=5, device=device)
        prompts = _SyntheticPrompts(cfg.text_embed_dim, device)
    else:
        base = Path(args.base_model)
        if not base.is_dir():
            raise RuntimeError(f"--base_model {args.base_model} is not a local diffusers directory (no network access)")
        tcfg = json.loads((base / "transformer" / "config.json").read_text())
        known = TransformerConfig.__dataclass_fields__.keys()
        cfg = TransformerConfig(**{k: v for k, v in tcfg.items() if k in known})
        transformer = CogVideoXTransformer3D(cfg, _load_safetensors_dir(base / "transformer"), device=device)
        vcfg = json.loads((base / "vae" / "config.json").read_text())
        vknown = VAEDecoderConfig.__dataclass_fields__.keys()
        vkw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in vcfg.items() if k in vknown}
        vae_sd = _load_safetensors_dir(base / "vae")
        vae = AutoencoderKLCogVideoXDecoder(vae_sd, VAEDecoderConfig(**vkw), device=device)
        prompts = _T5Prompts(base, device)
    vae.enable_tiling()
    vae.enable_slicing()
    # :18 `CogVideoXDPMScheduler.from_config(pipe.scheduler.config, timestep_spacing="trailing")`: betas / snr_shift_scale of the checkpoint
    scheduler = CogVideoXDPMScheduler() if args.synthetic else CogVideoXDPMScheduler.from_pretrained(args.base_model, timestep_spacing="trailing")
    pipe = CogVideoXDenoisePipeline(transformer, scheduler, vae=vae, vae_scaling_factor=vae.config.scaling_factor)
    pipe.vae_encoder = None
    if with_encoder:
        from ..vae import AutoencoderKLCogVideoXEncoder
        if args.synthetic:
            pipe.vae_encoder = AutoencoderKLCogVideoXEncoder.random_init(VAEDecoderConfig(), seed=6, device=device)
        else:
            pipe.vae_encoder = AutoencoderKLCogVideoXEncoder(vae_sd, VAEDecoderConfig(**vkw), device=device)
        pipe.vae_encoder.enable_tiling()                            # pipe.vae.enable_tiling() covers encode in the library
        pipe.vae_encoder.enable_slicing()
    return pipe, prompts


def generate(args):
    device = torch.device(f"cuda:{args.gpu_id}")
    torch.cuda.set_device(device)
    print(f"Loading base model: {args.base_model}")
    pipe, prompts = build_pipeline(args, device)

    if args.lora_path:
        if not os.path.exists(args.lora_path):
            print(f"LoRA path not found: {args.lora_path}, using base model")
        else:
            from ..lora import merge_lora
            print(f"Mounting LoRA: {args.lora_path}")
            merge_lora(pipe.transformer, args.lora_path)
            print("LoRA merged.")
    pipe.transformer.eval()

    tasks = load_tasks(args.prompt_json, args.num_prompts)
    if tasks is None:
        print("Unsupported JSON format")
        return
    print(f"Generating {len(tasks)} prompts, seed={args.seed}")
    output_root = Path(args.output_dir)
    output_root.mkdir(parents=True, exist_ok=True)
    negative = prompts("")

    for idx, item in enumerate(tasks):
        text_prompt = item.get("text_prompt", item.get("prompt", "")).strip()
        if not text_prompt:
            continue
        group_id, video_path = video_path_for(output_root, item, idx, args.seed)
        video_path.parent.mkdir(parents=True, exist_ok=True)
        if video_path.exists():
            print(f"[{idx+1}/{len(tasks)}] Skip existing: {group_id}")
            continue
        print(f"[{idx+1}/{len(tasks)}] Generating: {group_id}")
        try:
            generator = torch.Generator(device=device).manual_seed(args.seed)
            frames = pipe(prompts(text_prompt), negative, num_frames=args.num_frames, height=args.height, width=args.width,
                          num_inference_steps=args.num_inference_steps, guidance_scale=args.guidance_scale,
                          generator=generator, output_type="pt")
            export_to_video(frames[0], str(video_path), fps=args.fps)
        except Exception as e:                      # noqa: BLE001 — the reference continues with the next prompt
            print(f"  Failed: {e}")
        torch.cuda.empty_cache()
    print("Done.")


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="CogVideoX-5B T2V generation")
    parser.add_argument("--base_model", type=str, default="THUDM/CogVideoX-5B")
    parser.add_argument("--prompt_json", type=str, required=True)
    parser.add_argument("--output_dir", type=str, required=True)
    parser.add_argument("--lora_path", type=str, default=None, help="e.g. checkpoints/VideoGPA-T2V-lora")
    parser.add_argument("--gpu_id", type=int, default=0)
    parser.add_argument("--seed", type=int, default=42)
    parser.add_argument("--num_prompts", type=int, default=None)
    parser.add_argument("--num_inference_steps", type=int, default=50)
    parser.add_argument("--guidance_scale", type=float, default=6.0)
    parser.add_argument("--fps", type=int, default=8)
    # not in the reference CLI (the diffusers pipeline defaults): kept overridable for smoke runs
    parser.add_argument("--num_frames", type=int, default=49)
    parser.add_argument("--height", type=int, default=480)
    parser.add_argument("--width", type=int, default=720)
    parser.add_argument("--synthetic", type=int, default=0, metavar="N_LAYERS",
                        help="random-init N-block model + hash-seeded prompt embeddings (no checkpoint needed)")
    return parser


def main(argv=None):
    generate(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
