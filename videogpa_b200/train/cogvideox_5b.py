"""`train/CogVideoX-5B/03_train.py` of the reference without Lightning: the same configuration keys, data split, optimizer,
learning-rate schedule, gradient accumulation / clipping, validation and LoRA export around the sm_100a training step.

Reference: train/CogVideoX-5B/03_train.py:39-81 (DEFAULT_CONFIG; YAML `training:` override, `--config --devices --base_path`),
:207-214 (AdamW lr 5e-6 + `get_cosine_schedule_with_warmup(warmup_steps, max_steps)`), :226-248 (DPODataset, 98 % / 2 % split
with generator seed 42, batch_size 1), :252-280 (DDP, bf16, max_steps, accumulate_grad_batches 2, gradient_clip_val 1.0,
limit_val_batches 50, validation every epoch), :287 (`save_pretrained(out / "final_lora")`).

One process per GPU (launch with torchrun for `devices` > 1): every rank runs its shard of the pairs and the LoRA gradients
are averaged over NCCL (parallel.average_gradients). Logging is stdout (wandb is outside). A checkpoint whose transformer has
in_channels 32 (CogVideoX-5B-I2V, `train/CogVideoX-I2V-5B/03_train.py`) gets the VAE encoder for the image condition. `--synthetic N` swaps the 5B
checkpoint for an N-block random-weight transformer so the loop can be exercised without weights.
"""
from __future__ import annotations

import argparse
import math
import os
from pathlib import Path

import torch
from torch.utils.data import DataLoader, random_split

DEFAULT_CONFIG = {
    "devices": [0, 1, 2, 3, 4, 5, 6, 7],
    "metadata_path": "your_meta_data_t2v.json",
    "model_path": "THUDM/CogVideoX-5b",
    "output_dir": "your/outputs/root",
    "base_path": "/path/to/dataset",
    # DPO dataset
    "metric_name": "consistency_score", "metric_mode": "min", "min_gap": 0.05, "metric_threshold": 0.8, "motion_threshold": 0.001,
    # training
    "learning_rate": 5e-6, "beta": 1.0, "max_epochs": 100, "max_steps": 10000, "warmup_steps": 500, "batch_size": 1,
    "accumulate_grad_batches": 2, "gradient_clip_val": 1.0,
    # LoRA
    "lora_rank": 64, "lora_alpha": 128.0, "lora_dropout": 0.0, "lora_target_modules": ["to_q", "to_k", "to_v", "to_out.0"],
    # logging / checkpointing
    "experiment_name": "cogvideo_dpo_t2v", "checkpoint_every_n_steps": 1000, "log_every_n_steps": 10, "save_top_k": 10,
    # switches
    "enable_gradient_checkpointing": True, "enable_slicing": True, "enable_tiling": True,
}


def cosine_schedule_with_warmup(step: int, warmup_steps: int, total_steps: int) -> float:
    """LR multiplier of diffusers' get_cosine_schedule_with_warmup (num_cycles 0.5): linear warm-up, then half a cosine."""
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    progress = float(step - warmup_steps) / float(max(1, total_steps - warmup_steps))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * 2.0 * 0.5 * progress)))


def split_dataset(ds, seed: int = 42):
    """98 % train / 2 % validation, `random_split(..., generator=torch.Generator().manual_seed(42))` (:237-242)."""
    n_train = int(0.98 * len(ds))
    return random_split(ds, [n_train, len(ds) - n_train], generator=torch.Generator().manual_seed(seed))


def fit(step, train_loader, config: dict, val_loader=None, log=print, rank: int = 0, checkpoint_fn=None) -> dict:
    """The loop `trainer.fit` runs around training_step: gradient accumulation, clipping, AdamW + cosine warm-up schedule per
    optimizer step, DDP gradient average, validation (at most 50 batches) after every epoch. -> {"steps", "last_loss", "val"}.

    Stopping rule: the reference passes only `max_steps` to pl.Trainer (:258-266), so Lightning sets max_epochs = -1 and
    training runs until `max_steps` optimizer steps; `max_epochs` of the config is not a stopping criterion there and is
    ignored here too. Like Lightning, a partial accumulation window is flushed (stepped) on the last batch of an epoch; every
    rank must see the same number of micro-batches per epoch (main_train pads the shards like DistributedSampler), so all
    ranks issue the same sequence of all-reduces. `checkpoint_fn(gstep, val_loss)` is called on rank 0 every
    `checkpoint_every_n_steps` optimizer steps (ModelCheckpoint(every_n_train_steps=...), :267-274)."""
    import time
    import torch.distributed as dist
    from ..parallel import BucketedGradReducer
    opt = step.configure_optimizers(lr=config["learning_rate"])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: cosine_schedule_with_warmup(s, config.get("warmup_steps", 500), config["max_steps"]))
    params = step.trainable.parameters()
    acc, clip = int(config.get("accumulate_grad_batches", 1)), config.get("gradient_clip_val")
    max_steps = int(config["max_steps"])
    ckpt_every = int(config.get("checkpoint_every_n_steps", 0) or 0)
    ddp = dist.is_available() and dist.is_initialized()
    # DDP: bucketed all-reduce of the LoRA gradients, launched from autograd hooks while backward is still running
    reducer = BucketedGradReducer(params, bucket_bytes=int(config.get("ddp_bucket_bytes", 32 << 20))) if ddp else None
    gstep, last, val_hist, epoch = 0, float("nan"), [], 0
    t_start, world = time.time(), (dist.get_world_size() if ddp else 1)
    opt.zero_grad(set_to_none=True)

    def optimizer_step():
        nonlocal gstep
        if reducer is not None:
            reducer.finish()
        if clip:
            torch.nn.utils.clip_grad_norm_(params, float(clip))
        opt.step()
        sched.step()
        opt.zero_grad(set_to_none=True)
        gstep += 1
        if rank == 0 and gstep % int(config.get("log_every_n_steps", 10)) == 0:
            out = step.last_output
            # the scalars the reference logs (:169-186): loss, reward margin / accuracy, stats/samples_per_sec (global_step x
            # devices x batch_size over wall time) and stats/max_memory_gb
            sps = gstep * world * int(config["batch_size"]) / max(1e-9, time.time() - t_start)
            log(f"step {gstep}: train/loss {last:.5f} train/reward_margin {float(out.reward_margin):.5f} "
                f"train/reward_accuracy {float((out.reward_margin > 0).float().mean()):.2f} lr {sched.get_last_lr()[0]:.3e} "
                f"stats/samples_per_sec {sps:.4f} stats/max_memory_gb {(torch.cuda.max_memory_reserved() if torch.cuda.is_available() else 0) / 1024 ** 3:.1f}")
        if rank == 0 and checkpoint_fn is not None and ckpt_every and gstep % ckpt_every == 0:
            checkpoint_fn(gstep, val_hist[-1] if val_hist else float("nan"))

    while gstep < max_steps:
        n_batches = len(train_loader)
        if n_batches == 0:
            raise RuntimeError("fit: the training loader is empty")
        for bi, batch in enumerate(train_loader):
            window_end = (bi + 1) % acc == 0 or bi + 1 == n_batches   # windows restart every epoch; the last one may be short
            loss = step.training_step(batch)
            if reducer is not None:
                reducer.armed = window_end              # DDP's no_sync(): only the last micro-batch of a window reduces
            (loss / acc).backward()                     # Lightning divides the loss by accumulate_grad_batches
            last = float(loss.detach())
            if window_end:
                optimizer_step()
                if gstep >= max_steps:
                    break
        if val_loader is not None:
            tot, n = 0.0, 0
            for k, vb in enumerate(val_loader):
                if k >= 50:                             # limit_val_batches=50
                    break
                tot += float(step.validation_step(vb)["val/loss"])
                n += 1
            if n:
                val_hist.append(tot / n)
                if rank == 0:
                    log(f"epoch {epoch}: val/loss {tot / n:.5f}")
        epoch += 1
    if reducer is not None:
        reducer.remove()
    return {"steps": gstep, "last_loss": last, "val": val_hist, "epochs": epoch}


class TopKCheckpoints:
    """ModelCheckpoint(monitor="val/loss", mode="min", save_top_k=k, every_n_train_steps=n) of the reference (:267-274):
    every call writes `<dir>/step=<N>-val_loss=<v>` with `save_fn` and keeps the k entries with the lowest monitored value
    (entries saved before the first validation carry NaN and are the first to go)."""

    def __init__(self, directory, save_fn, top_k: int = 10):
        self.dir, self.save_fn, self.top_k, self.kept = Path(directory), save_fn, int(top_k), []

    def __call__(self, gstep: int, val_loss: float):
        import shutil
        path = self.dir / (f"step={gstep}-val_loss={val_loss:.4f}" if val_loss == val_loss else f"step={gstep}")
        self.save_fn(str(path))
        self.kept.append((val_loss if val_loss == val_loss else float("inf"), gstep, path))
        if self.top_k >= 0 and len(self.kept) > self.top_k:
            self.kept.sort(key=lambda e: (e[0], -e[1]))
            for _, _, old in self.kept[self.top_k:]:
                shutil.rmtree(old, ignore_errors=True)
            self.kept = self.kept[:self.top_k]
        return path


def main_train(config: dict, synthetic_layers: int = 0) -> dict:
    from ..dataset import DPODataset, collate_fn
    from ..parallel import init_from_env, shard_padded
    from ..train_dit import LoRATrainableTransformer
    from ..train_step import DPOSharedStep
    from ..transformer import CogVideoXTransformer3D, TransformerConfig
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = 0, 0
    if world > 1:
        rank, world, local = init_from_env("nccl")
    gpu = config["devices"][local] if local < len(config["devices"]) else local
    device = torch.device(f"cuda:{gpu}")
    torch.cuda.set_device(device)
    out_p = Path(config["output_dir"])
    (out_p / "checkpoints").mkdir(parents=True, exist_ok=True)

    full = DPODataset(base_path=config["base_path"], metadata_path=config["metadata_path"],
                      metric_name=config.get("metric_name", "consistency_score"), metric_mode=config.get("metric_mode", "min"),
                      min_gap=config.get("min_gap", 0.05), motion_threshold=config.get("motion_threshold", 0.001))
    train_ds, val_ds = split_dataset(full)
    if world > 1:                                       # DistributedSampler's job: equal-sized shards, padded by wrapping around
        train_ds = torch.utils.data.Subset(train_ds, shard_padded(range(len(train_ds)), rank, world))
    train_loader = DataLoader(train_ds, batch_size=config["batch_size"], shuffle=True, collate_fn=collate_fn,
                              generator=torch.Generator().manual_seed(1234 + rank))
    val_loader = DataLoader(val_ds, batch_size=1, shuffle=False, collate_fn=collate_fn) if len(val_ds) else None

    if synthetic_layers:
        variant = config.get("synthetic_variant")
        cfg = (TransformerConfig.cogvideox_5b_i2v() if variant == "i2v" else
               TransformerConfig.cogvideox1_5_5b() if variant == "1.5" else TransformerConfig.cogvideox_5b())
        cfg.num_layers = synthetic_layers
        transformer = CogVideoXTransformer3D.random_init(cfg, seed=1234, device=device)
    else:
        import json
        from ..generate.cogvideox_5b import _load_safetensors_dir
        base = Path(config["model_path"])
        if not base.is_dir():
            raise RuntimeError(f"model_path {config['model_path']} is not a local diffusers directory (no network access)")
        tcfg = json.loads((base / "transformer" / "config.json").read_text())
        known = TransformerConfig.__dataclass_fields__.keys()
        transformer = CogVideoXTransformer3D(TransformerConfig(**{k: v for k, v in tcfg.items() if k in known}),
                                             _load_safetensors_dir(base / "transformer"), device=device)
    if sorted(config["lora_target_modules"]) != sorted(["to_q", "to_k", "to_v", "to_out.0"]) or config.get("lora_dropout", 0.0) != 0.0:
        raise RuntimeError("only the reference's LoRA setup is supported: to_q / to_k / to_v / to_out.0, dropout 0")
    pol = LoRATrainableTransformer(transformer, r=config["lora_rank"], lora_alpha=config["lora_alpha"],
                                   gradient_checkpointing=bool(config.get("enable_gradient_checkpointing", True)))
    vae_encoder = None
    if transformer.config.in_channels == 32:            # CogVideoX-5B-I2V (train/CogVideoX-I2V-5B/03_train.py): image condition
        from ..vae import AutoencoderKLCogVideoXEncoder, VAEDecoderConfig
        if synthetic_layers:
            vae_encoder = AutoencoderKLCogVideoXEncoder.random_init(VAEDecoderConfig(), seed=6, device=device)
        else:
            import json
            from ..generate.cogvideox_5b import _load_safetensors_dir
            vcfg = json.loads((Path(config["model_path"]) / "vae" / "config.json").read_text())
            vknown = VAEDecoderConfig.__dataclass_fields__.keys()
            vkw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in vcfg.items() if k in vknown}
            vae_encoder = AutoencoderKLCogVideoXEncoder(_load_safetensors_dir(Path(config["model_path"]) / "vae"), VAEDecoderConfig(**vkw), device=device)
        if config.get("enable_slicing"):
            vae_encoder.enable_slicing()
        if config.get("enable_tiling"):
            vae_encoder.enable_tiling()
    scheduler = None                                    # :113 CogVideoXDPMScheduler.from_pretrained(model_path, subfolder="scheduler")
    if not synthetic_layers:
        from ..schedulers import CogVideoXDPMScheduler
        scheduler = CogVideoXDPMScheduler.from_pretrained(config["model_path"], subfolder="scheduler",
                                                          timestep_spacing="trailing")     # add_noise / get_velocity only read alphas_cumprod
    step = DPOSharedStep(transformer, None, beta=config["beta"], scheduler=scheduler, trainable=pol, vae_encoder=vae_encoder)
    ckpt = TopKCheckpoints(out_p / "checkpoints", pol.save_pretrained, top_k=int(config.get("save_top_k", 10))) if rank == 0 else None
    res = fit(step, train_loader, config, val_loader=val_loader, rank=rank, checkpoint_fn=ckpt)
    if rank == 0:
        pol.save_pretrained(str(out_p / "final_lora"))
        print(f"saved {out_p / 'final_lora'} after {res['steps']} optimizer steps")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return res


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument("--config", type=str, default=None)
    p.add_argument("--devices", type=str, default=None)
    p.add_argument("--base_path", type=str, default="/path/to/dataset")
    p.add_argument("--synthetic", type=int, default=0, help="N > 0: N-block random-weight transformer instead of --model_path (not a reference flag)")
    return p


def load_config(args) -> dict:
    """DEFAULT_CONFIG <- YAML `training:` section <- --devices (reference :290-305)."""
    config = dict(DEFAULT_CONFIG)
    config["base_path"] = args.base_path
    if args.config:
        import yaml
        with open(args.config, "r") as f:
            config.update(yaml.safe_load(f).get("training", {}))
    if args.devices:
        config["devices"] = [int(d) for d in args.devices.split(",")]
    return config


def main(argv=None):
    args = build_parser().parse_args(argv)
    return main_train(load_config(args), synthetic_layers=args.synthetic)


if __name__ == "__main__":
    main()
