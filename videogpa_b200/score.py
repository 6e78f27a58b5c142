"""`replicate_scorer.py` of the reference around the sm_100a scorer kernels: walk `<base_dir>/<prompt_id>/*.mp4`, score every
video with `VideoProcessor.process`, write the CSV / JSON report.

Reference: replicate_scorer.py:21-60 (the `SCORE_*` environment configuration, same names and defaults), :140-174 (task collection:
sorted prompt directories, sorted `*.mp4`, `SCORE_SEED_FILTER`, `SCORE_MAX_VIDEOS`), :177-188 (`SCORE_RESUME` from the JSON report),
:77-137 (per-video item: metric columns, `error` + `None` metrics when a video fails), :238-260 (contiguous chunks of
ceil(n / num_gpus) videos per GPU), :262-301 (merge, sort by (prompt_id, video_name), CSV, JSON `{config, items, summary}`).

What differs: one process per GPU under torchrun (`RANK` / `WORLD_SIZE`) instead of a `multiprocessing` pool, the per-rank result lists
are gathered with `all_gather_object`; the depth backbone (VGGT / Depth-Anything-3) and LPIPS / PSNR / SSIM (piq, lpips) are third-party
networks outside this build, so `backbone_fn(frames) -> predictions` is injected, `Consistency_Score` is built only when an LPIPS
callable is given and extra metrics can be passed in. Columns without a metric read 0.0, like `res.get(name, 0.0)` in the reference.
"""
from __future__ import annotations

import json
import os
from pathlib import Path

DEFAULT_VGGT_MODEL = "facebook/VGGT-1B"
DEFAULT_DA3_MODEL = "depth-anything/DA3-Large"
METRIC_COLS = ["psnr", "ssim", "lpips", "mvcs", "consistency_score", "epipolar"]


def parse_int_list_env(name, default):
    raw = os.getenv(name)
    if not raw:
        return list(default)
    return [int(item.strip()) for item in raw.split(",") if item.strip()]


def parse_bool_env(name, default):
    raw = os.getenv(name)
    if raw is None:
        return default
    return raw.strip().lower() in {"1", "true", "yes", "y", "on"}


def build_score_config() -> dict:
    """replicate_scorer.py:36-56."""
    backbone = os.getenv("SCORE_BACKBONE", "da3").strip().lower()
    default_model = DEFAULT_DA3_MODEL if backbone == "da3" else DEFAULT_VGGT_MODEL
    return {
        "devices": parse_int_list_env("SCORE_DEVICES", [0]),
        "base_dir": os.getenv("SCORE_BASE_DIR", "output/replicate"),
        "output_csv": os.getenv("SCORE_OUTPUT_CSV", "output/replicate/scores.csv"),
        "output_json": os.getenv("SCORE_OUTPUT_JSON", ""),
        "num_frames": int(os.getenv("SCORE_NUM_FRAMES", "10")),
        "conf_thres": int(os.getenv("SCORE_CONF_THRES", "0")),
        "ignore_seed": parse_bool_env("SCORE_IGNORE_SEED", True),
        "descriptor_type": os.getenv("SCORE_DESCRIPTOR_TYPE", "lightglue"),
        "backbone": backbone,
        "model_name": os.getenv("SCORE_MODEL_NAME", default_model),
        "resume": parse_bool_env("SCORE_RESUME", False),
        "max_videos": int(os.getenv("SCORE_MAX_VIDEOS", "0")),
        "seed_filter": os.getenv("SCORE_SEED_FILTER", ""),
    }


def build_metrics(device, config: dict, lpips_net=None, extra: dict | None = None) -> dict:
    """replicate_scorer.py:63-74 with what this build provides: MSE, MVCS, Epipolar, and Consistency_Score when an LPIPS callable is
    injected. `extra` = {"PSNR": metric, "SSIM": metric, "LPIPS": metric, ...} objects with `compute(gt=, rep=)`."""
    from .metrics import Consistency_Score, EpipolarMetric, MSEMetric, MVCSMetric
    metrics = {"MSE": MSEMetric()}
    if lpips_net is not None:
        metrics["Consistency_Score"] = Consistency_Score(lpips_net, device=device)
    metrics["MVCS"] = MVCSMetric(device=device)
    metrics["Epipolar"] = EpipolarMetric(descriptor_type=config["descriptor_type"], device=device)
    metrics.update(extra or {})
    return metrics


def collect_all_video_tasks(config: dict) -> list:
    """replicate_scorer.py:140-174."""
    base_path = Path(config["base_dir"])
    all_tasks = []
    if not base_path.exists():
        print(f"Base dir does not exist: {base_path}")
        return all_tasks
    print(f"Scanning benchmark root: {base_path}")
    for prompt_dir in sorted(base_path.iterdir()):
        if not prompt_dir.is_dir():
            continue
        for v_file in sorted(prompt_dir.glob("*.mp4")):
            seed_filter = config.get("seed_filter", "")
            if seed_filter and f"seed_{seed_filter}" not in v_file.name:
                continue
            all_tasks.append({"path": v_file, "prompt_id": prompt_dir.name, "relative_path": str(v_file.relative_to(base_path))})
    if config["max_videos"] > 0:
        all_tasks = all_tasks[: config["max_videos"]]
    return all_tasks


def load_existing_items(config: dict) -> dict:
    """replicate_scorer.py:177-188."""
    output_json = config.get("output_json")
    if not config["resume"] or not output_json:
        return {}
    out_path = Path(output_json)
    if not out_path.exists():
        return {}
    with open(out_path, "r", encoding="utf-8") as f:
        payload = json.load(f)
    return {item["relative_path"]: item for item in payload.get("items", [])}


def chunk_tasks(tasks: list, num_workers: int) -> list:
    """replicate_scorer.py:238-244: contiguous chunks of ceil(n / num_workers), padded with empty chunks."""
    chunk_size = (len(tasks) + num_workers - 1) // num_workers if tasks else 0
    chunks = [tasks[i:i + chunk_size] for i in range(0, len(tasks), chunk_size)] if chunk_size else []
    while len(chunks) < num_workers:
        chunks.append([])
    return chunks


def score_tasks(processor, tasks: list, config: dict, tag: str = "GPU-0") -> list:
    """The loop of `score_worker` (replicate_scorer.py:100-137) on an already built VideoProcessor."""
    scored = []
    for task in tasks:
        v_path = Path(task["path"])
        item = {"prompt_id": task["prompt_id"], "video_name": v_path.name, "video_path": str(v_path),
                "relative_path": task["relative_path"], "backbone": config["backbone"]}
        try:
            results = processor.process(video_path=str(v_path), thresholds=[config["conf_thres"]], num_frames=config["num_frames"],
                                        save_visuals=False)
            res = results.get(config["conf_thres"], {})
            item.update({
                "mse": float(res.get("MSE")),                            # the reference has no default here: a missing MSE is an error
                "consistency_score": float(res.get("Consistency_Score", 0.0)),
                "motion_score": float(res.get("motion_norm", 0.0)),
                "psnr": float(res.get("PSNR", 0.0)),
                "ssim": float(res.get("SSIM", 0.0)),
                "lpips": float(res.get("LPIPS", 0.0)),
                "mvcs": float(res.get("MVCS", 0.0)),
                "epipolar": float(res.get("Epipolar", 0.0)),
            })
        except Exception as exc:                                         # noqa: BLE001  (reference :129-133)
            print(f"\nWarning {tag} failed on {v_path.name}: {exc}")
            item["error"] = str(exc)
            for metric_name in METRIC_COLS:
                item.setdefault(metric_name, None)
        scored.append(item)
    return scored


def build_summary(df) -> dict:
    """replicate_scorer.py:191-208."""
    import pandas as pd
    summary = {"overall": {"video_count": int(len(df))}}
    if df.empty:
        return summary
    for metric in METRIC_COLS:
        if metric not in df.columns:
            df[metric] = None
    summary["overall"].update({metric: (None if pd.isna(df[metric].mean()) else float(df[metric].mean())) for metric in METRIC_COLS})
    return summary


def write_reports(flat_results: list, config: dict):
    """CSV + JSON (replicate_scorer.py:211-228, 276-283). -> the DataFrame."""
    import pandas as pd
    df = pd.DataFrame(flat_results)
    if config.get("output_csv"):
        out_file = Path(config["output_csv"])
        out_file.parent.mkdir(parents=True, exist_ok=True)
        df.to_csv(out_file, index=False, encoding="utf-8")
        print(f"\nSaved CSV report to: {out_file}")
    if config.get("output_json"):
        payload = {"config": config, "items": flat_results, "summary": build_summary(df)}
        out_path = Path(config["output_json"])
        out_path.parent.mkdir(parents=True, exist_ok=True)
        with open(out_path, "w", encoding="utf-8") as f:
            json.dump(payload, f, indent=2, ensure_ascii=False)
        print(f"\nSaved JSON report to: {out_path}")
    return df


def main(backbone_fn=None, lpips_net=None, extra_metrics: dict | None = None, processor_factory=None, config: dict | None = None):
    """`python -m videogpa_b200.score` needs a backbone: call `main(backbone_fn=...)` from a script that owns the VGGT / DA3 model.
    `processor_factory(config, device) -> object with .process(...)` replaces the VideoProcessor construction (tests)."""
    config = config or build_score_config()
    all_tasks = collect_all_video_tasks(config)
    if not all_tasks:
        print("No videos found for scoring.")
        return None
    existing_items = load_existing_items(config)
    pending = [t for t in all_tasks if t["relative_path"] not in existing_items]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = 0, 0
    if world > 1:
        from .parallel import init_from_env
        rank, world, local = init_from_env()
    if rank == 0:
        print("\nScoring config")
        print(f"  Backbone    : {config['backbone']}")
        print(f"  Model       : {config['model_name']}")
        print(f"  Devices     : {config['devices']}")
        print(f"  Total videos: {len(all_tasks)}")
        print(f"  Pending     : {len(pending)}")
    new_results = []
    if pending:
        mine = chunk_tasks(pending, world)[rank]
        gpu = config["devices"][local] if local < len(config["devices"]) else local
        if processor_factory is not None:
            processor = processor_factory(config, gpu)
        else:
            import torch
            from .process_video import VideoProcessor
            if backbone_fn is None:
                raise RuntimeError("score.main needs backbone_fn(frames) -> predictions: the VGGT / DA3 backbones are outside this build")
            device = torch.device(f"cuda:{gpu}")
            torch.cuda.set_device(device)
            processor = VideoProcessor(metrics=build_metrics(device, config, lpips_net, extra_metrics), model_name=config["model_name"],
                                       device=device, backbone=config["backbone"], backbone_fn=backbone_fn)
        new_results = score_tasks(processor, mine, config, tag=f"GPU-{gpu}")
        if world > 1:
            import torch.distributed as dist
            gathered = [None] * world
            dist.all_gather_object(gathered, new_results)
            new_results = [item for sub in gathered for item in sub]
    if rank != 0:
        return None
    merged = dict(existing_items)
    for item in new_results:
        merged[item["relative_path"]] = item
    flat_results = sorted(merged.values(), key=lambda item: (item["prompt_id"], item["video_name"]))
    if not flat_results:
        print("No valid scoring results were produced.")
        return None
    df = write_reports(flat_results, config)
    print("\n========================================")
    print("Overall Mean Metrics")
    print("========================================")
    for metric in METRIC_COLS:
        if metric not in df.columns:
            df[metric] = None
    print(df[METRIC_COLS].apply(lambda c: c.astype(float)).mean().to_frame().T.rename(index={0: "overall"}).round(4))
    print(f"\nTotal videos scored: {len(df)}")
    print("\nAll scoring tasks completed.")
    return df


if __name__ == "__main__":
    main()
