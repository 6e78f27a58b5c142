"""`generate/Wan2.2-TI2V-5B.py` of the reference on the sm_100a Wan DiT — same flags, task rules and output layout.

Reference: generate/Wan2.2-TI2V-5B.py:43-157. Kept: `--model_path(req) --prompt_json(req) --output_dir(req) --lora_path
--lora_weight(0.2) --base_dir --gpu_id --seed(42) --num_prompts --frame_num(81) --shift(5.0) --sampling_steps(50)
--guide_scale(5.0) --fps(24)`; tasks = dict items or list entries (`group_id` or the index), items without prompt or image
skipped, `Image not found ... skipping`, `<output_dir>/<group_id>/seed_<seed>.*`, skip-if-exists, per-prompt
`try/except -> "Failed"`, LoRA `scaling *= lora_weight` before the merge (:66-70).

What runs here: the guided denoise loop of `WanTI2V.generate` on videogpa_b200 kernels — 30-block DiT with per-token
timesteps (first latent frame t = 0 and clamped to the encoded image), cond / uncond forwards, `uncond + g (cond - uncond)`,
shifted flow-matching schedule (`--shift`) and the UniPC multistep sampler that `WanTI2V.generate` defaults to
(schedulers.FlowUniPCMultistepScheduler; `sample_solver="euler"` selects the fused Euler step). The un-vendored Wan2.2
repository's Wan-VAE is outside this build (and its umT5 text encoder runs here only when the checkpoint directory carries it:
`WanPromptEncoder`), so the result of a prompt is the final latent
`seed_<seed>.latents.pt` ([48, F, h, w], bf16) instead of an mp4, and inputs are either synthetic (`--synthetic N`: N-block
random DiT, context / first-frame latent seeded from the prompt and image bytes) or precomputed next to the image:
`<image>.context.pt` ([L <= 512, 4096]), `<image>.latent.pt` ([48, 1, h, w]) and the encoded negative prompt
`<image>.null_context.pt` (or one shared `<model_path>/null_context.pt`) for the unconditional branch.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
from pathlib import Path

import torch

BF16 = torch.bfloat16


def load_tasks(prompt_json: str, num_prompts: int | None):
    """generate/Wan2.2-TI2V-5B.py:77-89."""
    with open(prompt_json, "r", encoding="utf-8") as f:
        raw = json.load(f)
    if isinstance(raw, dict):
        tasks = list(raw.items())
    elif isinstance(raw, list):
        tasks = [(item.get("group_id", i), item) for i, item in enumerate(raw)]
    else:
        return None
    return tasks[:num_prompts] if num_prompts else tasks


def latent_grid(frame_num: int, height: int, width: int):
    """Wan2.2 VAE strides (4, 16, 16): (F, h, w) of the latent and the token count after 1x2x2 patching."""
    F_ = (frame_num - 1) // 4 + 1
    h, w = height // 16, width // 16
    return F_, h, w, F_ * (h // 2) * (w // 2)


class WanTI2VEngine:
    """The part of `WanTI2V.generate(input_prompt, img, frame_num, shift, sampling_steps, guide_scale, seed)` that runs
    on the DiT: conditioning tensors in, final latent out."""

    def __init__(self, model, device):
        from ..wan import WanDenoiseStep
        self.model, self.device = model, device
        self._step = WanDenoiseStep

    @torch.no_grad()
    def generate(self, context, first_latent, frame_num=81, shift=5.0, sampling_steps=50, guide_scale=5.0, seed=42, size=(704, 1280),
                 sample_solver="unipc", null_context=None):
        """null_context: the umT5 encoding of the configured negative prompt (`sample_neg_prompt`), which `WanTI2V.generate`
        feeds to the unconditional branch; None = a single all-zero token (only meaningful for synthetic runs)."""
        from ..wan import flow_sigmas
        F_, h, w, S = latent_grid(frame_num, size[0], size[1])
        if tuple(first_latent.shape) != (self.model.config.in_dim, 1, h, w):
            raise RuntimeError(f"first-frame latent must be {(self.model.config.in_dim, 1, h, w)}, got {tuple(first_latent.shape)}")
        g = torch.Generator(device=self.device).manual_seed(seed)
        lat = torch.randn(self.model.config.in_dim, F_, h, w, device=self.device, generator=g).to(BF16)
        first = first_latent.to(device=self.device, dtype=BF16)
        lat[:, :1] = first
        hw = (h // 2) * (w // 2)
        step = self._step(self.model, guide_scale=guide_scale)
        if null_context is None:
            null = torch.zeros(1, context.shape[-1], device=self.device, dtype=BF16)
        else:
            null = null_context.to(device=self.device, dtype=BF16)
        ctx = context.to(device=self.device, dtype=BF16)
        if sample_solver == "unipc":                                   # WanTI2V.generate's default sampler
            from ..schedulers import FlowUniPCMultistepScheduler
            sch = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1.0)
            sch.set_timesteps(sampling_steps, shift=shift)
            x = lat.float()
            for i in range(sampling_steps):
                t = torch.full((1, S), float(int(sch.timesteps[i])))    # Wan's FlowUniPC scheduler hands the model int64 timesteps
                t[:, :hw] = 0                                          # the image frame is clean: t = 0 for its tokens
                v = step.guided_velocity(x, t, ctx, null)
                x = sch.step(v, x)
                x[:, :1] = first.float()                               # TI2V: the first latent frame stays the encoded image
            return x.to(BF16)
        if sample_solver != "euler":
            raise RuntimeError(f"unknown sample_solver {sample_solver!r} (unipc | euler)")
        sig = flow_sigmas(sampling_steps, shift)
        for i in range(sampling_steps):
            t = torch.full((1, S), sig[i] * 1000.0)
            t[:, :hw] = 0
            lat = step(lat, t, sig[i], sig[i + 1], ctx, null, first_frame=first)
        return lat


class WanPromptEncoder:
    """`self.text_encoder([prompt], device)` of `WanTI2V.generate`: umT5-XXL over the prompt padded to `text_len` = 512 with a mask,
    output trimmed to the true length -> context [L, 4096]. Runs on videogpa_b200.t5 (umT5 = one relative-position table per block).
    Loaded only when the checkpoint directory carries the encoder: `models_t5_umt5-xxl-enc-bf16.pth` (Wan's parameter names, mapped by
    t5.wan_umt5_to_transformers_names) or `text_encoder/` in transformers' layout, plus the tokenizer under `google/umt5-xxl/`."""

    TEXT_LEN = 512

    def __init__(self, model_path: Path, device):
        from transformers import AutoTokenizer
        from ..t5 import T5Config, T5EncoderModel, wan_umt5_to_transformers_names
        pth = model_path / "models_t5_umt5-xxl-enc-bf16.pth"
        if pth.is_file():
            sd = wan_umt5_to_transformers_names(torch.load(str(pth), map_location="cpu"))
        elif (model_path / "text_encoder").is_dir():
            from .cogvideox_5b import _load_safetensors_dir
            sd = _load_safetensors_dir(model_path / "text_encoder")
        else:
            raise RuntimeError("no umT5 weights in the checkpoint directory")
        self.tok = AutoTokenizer.from_pretrained(str(model_path / "google" / "umt5-xxl"))
        self.enc = T5EncoderModel(T5Config.umt5_xxl(), sd, device=device)
        self.device = device

    @staticmethod
    def available(model_path: Path) -> bool:
        return ((model_path / "models_t5_umt5-xxl-enc-bf16.pth").is_file() or (model_path / "text_encoder").is_dir()) and \
            (model_path / "google" / "umt5-xxl").is_dir()

    @torch.no_grad()
    def __call__(self, prompt: str) -> torch.Tensor:
        text = " ".join(prompt.split())                              # whitespace clean-up of Wan's tokenizer wrapper
        t = self.tok([text], padding="max_length", truncation=True, max_length=self.TEXT_LEN, add_special_tokens=True, return_tensors="pt")
        ids, mask = t.input_ids.to(self.device), t.attention_mask.to(self.device)
        L = int(mask.sum())
        return self.enc(ids, attention_mask=mask)[0][0, :L].cpu()


def _seeded(path_or_text, shape, salt: str):
    data = Path(path_or_text).read_bytes() if salt == "img" else path_or_text.encode()
    seed = int.from_bytes(hashlib.sha256(data).digest()[:4], "little")
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Wan2.2 TI2V generation")
    p.add_argument("--model_path", type=str, required=True, help="Wan2.2-TI2V-5B model path")
    p.add_argument("--prompt_json", type=str, required=True)
    p.add_argument("--output_dir", type=str, required=True)
    p.add_argument("--lora_path", type=str, default=None, help="Path to LoRA weights")
    p.add_argument("--lora_weight", type=float, default=0.2, help="LoRA strength relative to trained scaling (1.0 = full; default 0.2)")
    p.add_argument("--base_dir", type=str, default=None, help="Base dir for relative image paths")
    p.add_argument("--gpu_id", type=int, default=0)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--num_prompts", type=int, default=None)
    p.add_argument("--frame_num", type=int, default=81)
    p.add_argument("--shift", type=float, default=5.0)
    p.add_argument("--sampling_steps", type=int, default=50)
    p.add_argument("--guide_scale", type=float, default=5.0)
    p.add_argument("--fps", type=int, default=24)
    # additions of this build (not reference flags)
    p.add_argument("--synthetic", type=int, default=0, help="N > 0: N-block random-weight DiT and seeded conditioning (no checkpoint needed)")
    p.add_argument("--height", type=int, default=704)
    p.add_argument("--width", type=int, default=1280)
    return p


def generate(args):
    from ..wan import WanConfig, WanTransformer3D
    device = torch.device(f"cuda:{args.gpu_id}")
    torch.cuda.set_device(device)
    print(f"Loading Wan TI2V engine: {args.model_path}")
    cfg = WanConfig.ti2v_5b()
    if args.synthetic:
        cfg.num_layers = args.synthetic
        model = WanTransformer3D.random_init(cfg, seed=21, device=device)
    else:
        from .cogvideox_5b import _load_safetensors_dir
        mp = Path(args.model_path)
        if not mp.is_dir():
            raise RuntimeError(f"--model_path {args.model_path} is not a local directory (no network access)")
        model = WanTransformer3D(cfg, _load_safetensors_dir(mp), device=device)
    if args.lora_path:
        if not Path(args.lora_path).exists():
            print(f"LoRA path not found: {args.lora_path}, using base model")
        else:
            from ..lora import merge_lora
            print(f"Mounting LoRA: {args.lora_path} (weight={args.lora_weight})")
            merge_lora(model, args.lora_path, weight=args.lora_weight)
            print("LoRA merged.")
    engine = WanTI2VEngine(model, device)
    prompt_encoder = None
    if not args.synthetic and WanPromptEncoder.available(Path(args.model_path)):
        print("Loading umT5 prompt encoder")
        prompt_encoder = WanPromptEncoder(Path(args.model_path), device)
    tasks = load_tasks(args.prompt_json, args.num_prompts)
    if tasks is None:
        print("Unsupported JSON format")
        return
    print(f"Generating {len(tasks)} prompts, seed={args.seed}")
    output_root = Path(args.output_dir)
    output_root.mkdir(parents=True, exist_ok=True)
    F_, h, w, _ = latent_grid(args.frame_num, args.height, args.width)
    for idx, (group_id, item) in enumerate(tasks):
        group_id = str(group_id).replace("/", "_")
        text_prompt = item.get("text_prompt", item.get("prompt", "")).strip()
        image_path = item.get("image_prompt", item.get("image_path", ""))
        if not text_prompt or not image_path:
            continue
        if not Path(image_path).exists() and args.base_dir:
            image_path = str(Path(args.base_dir) / image_path)
        if not Path(image_path).exists():
            print(f"[{idx+1}/{len(tasks)}] Image not found: {image_path}, skipping")
            continue
        out_dir = output_root / group_id
        out_dir.mkdir(parents=True, exist_ok=True)
        out_path = out_dir / f"seed_{args.seed}.latents.pt"
        if out_path.exists():
            print(f"[{idx+1}/{len(tasks)}] Skip existing: {group_id}")
            continue
        print(f"[{idx+1}/{len(tasks)}] Generating: {group_id}")
        try:
            if args.synthetic:
                context = _seeded(text_prompt, (64, cfg.text_dim), "txt")
                first = _seeded(image_path, (cfg.in_dim, 1, h, w), "img")
            else:
                cpath, lpath = Path(str(image_path) + ".context.pt"), Path(str(image_path) + ".latent.pt")
                if not lpath.exists():
                    raise RuntimeError(f"{lpath.name} not found: the Wan-VAE (first-frame latent) is outside this build")
                first = torch.load(str(lpath), map_location="cpu")
                if cpath.exists():
                    context = torch.load(str(cpath), map_location="cpu")
                elif prompt_encoder is not None:
                    context = prompt_encoder(text_prompt)
                else:
                    raise RuntimeError(f"{cpath.name} not found and the checkpoint directory has no umT5 encoder")
                # the unconditional branch sees the encoded negative prompt: per image, else one shared file in the model directory,
                # else <model_path>/negative_prompt.txt (the configuration's sample_neg_prompt) through the umT5 encoder
                npath = Path(str(image_path) + ".null_context.pt")
                if not npath.exists():
                    npath = Path(args.model_path) / "null_context.pt"
                ntxt = Path(args.model_path) / "negative_prompt.txt"
                if npath.exists():
                    null_context = torch.load(str(npath), map_location="cpu")
                elif prompt_encoder is not None and ntxt.is_file():
                    null_context = prompt_encoder(ntxt.read_text(encoding="utf-8"))
                else:
                    raise RuntimeError(f"{Path(str(image_path)).name}.null_context.pt / {npath} / {ntxt.name} not found: the umT5 encoding of "
                                       "the negative prompt is needed for the unconditional branch")
            if args.synthetic:
                null_context = _seeded("negative prompt", (32, cfg.text_dim), "txt")
            lat = engine.generate(context, first, frame_num=args.frame_num, shift=args.shift, sampling_steps=args.sampling_steps,
                                  guide_scale=args.guide_scale, seed=args.seed, size=(args.height, args.width), null_context=null_context)
            torch.save(lat.cpu(), str(out_path))
        except Exception as e:                      # noqa: BLE001
            print(f"  Failed: {e}")
        torch.cuda.empty_cache()
    print("Done.")


def main(argv=None):
    generate(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
