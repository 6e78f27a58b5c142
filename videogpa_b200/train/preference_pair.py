"""`train/01_preference_pair.py` of the reference: score every video of every prompt group with the Consistency_Score metric and write
the metadata JSON that `train/dataset.py` (videogpa_b200.dataset.DPODataset) turns into preference pairs.

Reference: train/01_preference_pair.py:22-35 (GPUS, INPUT_JSON, OUTPUT_JSON, CONF_THRES = 0, NUM_FRAMES = 10), :43-73 (safe JSON read /
atomic write), :78-205 (worker: per-process de-duplication by video path, resume from already scored entries, file validation,
`vp.process(...)` -> `consistency_score` + `motion_norm` stored as floats, groups without videos dropped), :210-283 (input formats
`{"groups": [...]}` or a plain list, resume map from the existing output, round-robin split `all_groups[i::num_gpus]`, merge, save).

One process per GPU under torchrun instead of a spawn pool (rank r scores `all_groups[r::world]`, results gathered with
`all_gather_object`); the VGGT backbone and LPIPS are third-party, so a `processor` (anything with VideoProcessor's `process(...)`) or a
`backbone_fn` + `lpips_net` pair is injected.
"""
from __future__ import annotations

import json
import logging
import os
from pathlib import Path

CONF_THRES = 0
NUM_FRAMES = 10


def safe_load_json(path):
    """:43-56 — None when the file is missing or unreadable."""
    try:
        if not os.path.exists(path):
            logging.warning(f"JSON file not found: {path}")
            return None
        with open(path, "r", encoding="utf-8") as f:
            return json.load(f)
    except Exception as e:                                        # noqa: BLE001
        logging.error(f"Failed to load JSON file {path}: {str(e)}")
        return None


def safe_save_json(path, data) -> None:
    """:58-73 — write to `<path>.tmp`, then replace (no half-written file on interrupt)."""
    temp_path = str(path) + ".tmp"
    try:
        with open(temp_path, "w", encoding="utf-8") as f:
            json.dump(data, f, indent=4, ensure_ascii=False)
        os.replace(temp_path, path)
        logging.info(f"Successfully saved JSON to {path} (total items: {len(data) if isinstance(data, list) else 'object'})")
    except Exception as e:                                        # noqa: BLE001
        logging.error(f"Failed to save JSON file {path}: {str(e)}")
        if os.path.exists(temp_path):
            os.remove(temp_path)


def extract_groups(original_data):
    """:222-238 — `{"groups": [...]}` or a plain list; None for anything else."""
    if isinstance(original_data, dict):
        if isinstance(original_data.get("groups"), list):
            return original_data["groups"]
        logging.error("Input JSON is an object but missing 'groups' key or value is not a list.")
        return None
    if isinstance(original_data, list):
        return original_data
    logging.error("Input JSON format not supported.")
    return None


def build_scored_video_map(existing_output) -> dict:
    """:247-256 — videos of an earlier (plain-list) output that already carry both scores."""
    scored = {}
    if existing_output and isinstance(existing_output, list):
        for group in existing_output:
            for entry in group.get("videos", []):
                vp = entry.get("video_path")
                if vp and all(k in entry for k in ["consistency_score", "motion_norm"]):
                    scored[vp] = entry
    return scored


def score_groups(processor, groups_chunk: list, scored_video_map: dict, tag: str = "Worker-0", conf_thres: int = CONF_THRES,
                 num_frames: int = NUM_FRAMES) -> list:
    """The loop of `gpu_worker` (:112-205) on an already built processor."""
    seen = set()
    processed = []
    for group in groups_chunk:
        new_group = group.copy()
        videos = new_group.get("videos", [])
        if not isinstance(videos, list):
            videos = []
        scored_videos = []
        for video_entry in videos:
            entry = video_entry.copy()
            vpath = entry.get("video_path")
            if not vpath:                                         # entries without a path are kept as they are
                scored_videos.append(entry)
                continue
            if vpath in seen:                                     # in-process de-duplication: kept, not scored twice
                scored_videos.append(entry)
                continue
            seen.add(vpath)
            old = scored_video_map.get(vpath)
            if old and all(k in old for k in ["consistency_score", "motion_norm"]):        # resume
                entry.update({"consistency_score": old["consistency_score"], "motion_norm": old["motion_norm"]})
                scored_videos.append(entry)
                continue
            p = Path(vpath)
            if not p.exists():
                logging.warning(f"{tag}: Video not found - {vpath}")
                scored_videos.append(entry)
                continue
            if not os.access(p, os.R_OK):
                logging.warning(f"{tag}: No read permission - {vpath}")
                scored_videos.append(entry)
                continue
            if p.stat().st_size <= 0:
                logging.warning(f"{tag}: Empty video file - {vpath}")
                scored_videos.append(entry)
                continue
            try:
                results = processor.process(video_path=str(p), thresholds=[conf_thres], num_frames=num_frames, save_visuals=False, out_dir=None)
                res = results.get(conf_thres, {})
                cs, mn = res.get("Consistency_Score"), res.get("motion_norm")
                if cs is not None and mn is not None:
                    entry["consistency_score"] = float(cs)
                    entry["motion_norm"] = float(mn)
                else:
                    logging.warning(f"{tag}: No valid scores for - {vpath}")
            except Exception as e:                                # noqa: BLE001
                logging.warning(f"{tag}: Failed to process video {vpath}: {str(e)}")
            scored_videos.append(entry)
        new_group["videos"] = scored_videos
        if scored_videos:                                         # groups without any video are dropped
            processed.append(new_group)
    return processed


def process_video_scoring(input_json: str, output_json: str, processor=None, backbone_fn=None, lpips_net=None, devices=(0,),
                          conf_thres: int = CONF_THRES, num_frames: int = NUM_FRAMES):
    """:210-283. -> the list written to `output_json` (rank 0), None on the other ranks or when the input is unusable."""
    original = safe_load_json(input_json)
    if not original:
        logging.error("Task Terminated: Failed to load valid Input JSON.")
        return None
    all_groups = extract_groups(original)
    if all_groups is None:
        return None
    if len(all_groups) == 0:
        logging.warning("Extracted group list is empty, nothing to process.")
        return None
    scored_map = build_scored_video_map(safe_load_json(output_json))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = 0, 0
    if world > 1:
        from ..parallel import init_from_env
        rank, world, local = init_from_env()
    if processor is None:
        import torch
        from ..metrics import Consistency_Score
        from ..process_video import VideoProcessor
        if backbone_fn is None or lpips_net is None:
            raise RuntimeError("process_video_scoring needs `processor`, or `backbone_fn` and `lpips_net`: VGGT and LPIPS-VGG are outside this build")
        gpu = devices[local] if local < len(devices) else local
        device = torch.device(f"cuda:{gpu}")
        torch.cuda.set_device(device)
        processor = VideoProcessor(metrics={"Consistency_Score": Consistency_Score(lpips_net, device=device)}, model_name="facebook/VGGT-1B",
                                   device=device, backbone_fn=backbone_fn)
    mine = score_groups(processor, all_groups[rank::world], scored_map, tag=f"Worker-{rank}", conf_thres=conf_thres, num_frames=num_frames)
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        mine = [g for sub in gathered for g in sub]
    if rank != 0:
        return None
    safe_save_json(output_json, mine)
    logging.info(f"Stats: Input {len(original)} groups, Output {len(mine)} valid groups. Reused {len(scored_map)} previously scored videos.")
    return mine
