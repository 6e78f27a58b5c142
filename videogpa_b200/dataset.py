"""DPO preference-pair dataset — the on-disk formats and the pair-selection rule of `train/dataset.py`
(reference train/dataset.py:1-31 format, :102-201 selection, :206-284 loading / collate).

Pure host code (json + torch.load): it feeds the paired denoise forward + DPO loss of config 5 with
  * `meta_data.json`: {"groups": [{group_id, text_prompt | prompt, image_path | input_image_path, original_video_path,
    videos: [{video_path, generation_id, consistency_score, motion_norm, latent_path, condition_path}]}]}
  * latents `.pt` [16, 13, 60, 90] (CogVideoX) / [48, 21, h, w] (Wan), condition dicts with `encoder_hidden_states`
    [226, 4096] and optionally `image_embeds` (CogVideoX I2V) or `image_latent` (Wan TI2V).
Selection per group (lower metric is better for "min"): keep videos that have the metric, `motion_norm`, both paths, existing
files and `motion_norm >= motion_threshold`; need >= 2; sort by the metric (stable, descending for "max"); winner = first,
loser = last; drop the group if the winner misses `metric_threshold` or `|winner - loser| < min_gap`.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List, Optional

import torch
from torch.utils.data import Dataset


def _load(path: Path):
    try:
        return torch.load(path, weights_only=True)
    except TypeError:
        return torch.load(path)


def select_preference_pairs(groups: List[dict], base_path: Path, metric_name: str = "consistency_score", metric_mode: str = "min",
                            min_gap: float = 0.1, metric_threshold: Optional[float] = None,
                            motion_threshold: float = 0.001) -> List[Dict[str, Any]]:
    pairs = []
    for group in groups:
        videos = group.get("videos", [])
        if len(videos) < 2:
            continue
        valid = [v for v in videos
                 if metric_name in v and "motion_norm" in v and "latent_path" in v and "condition_path" in v
                 and (base_path / v["latent_path"]).exists() and (base_path / v["condition_path"]).exists()
                 and not (v["motion_norm"] < motion_threshold)]
        if len(valid) < 2:
            continue
        ranked = sorted(valid, key=lambda v: v[metric_name], reverse=(metric_mode == "max"))
        winner, loser = ranked[0], ranked[-1]
        wm, lm = winner[metric_name], loser[metric_name]
        if metric_threshold is not None:
            if (metric_mode == "min" and wm >= metric_threshold) or (metric_mode != "min" and wm <= metric_threshold):
                continue
        gap = abs(wm - lm)
        if gap < min_gap:
            continue
        pairs.append({"group_id": group.get("group_id", "unknown"), "prompt": group.get("text_prompt", group.get("prompt", "")),
                      "input_image_path": group.get("image_path", group.get("input_image_path")),
                      "original_video_path": group.get("original_video_path"), "winner": winner, "loser": loser, "metric_gap": gap})
    return pairs


class DPODataset(Dataset):
    def __init__(self, base_path: str, metadata_path: str, metric_name: str = "consistency_score", metric_mode: str = "min",
                 min_gap: float = 0.1, metric_threshold: Optional[float] = None, motion_threshold: float = 0.001,
                 max_samples: Optional[int] = None):
        super().__init__()
        self.base_path = Path(base_path)
        self.metadata_path = Path(metadata_path)
        self.metric_name, self.metric_mode = metric_name, metric_mode
        self.min_gap, self.metric_threshold, self.motion_threshold = min_gap, metric_threshold, motion_threshold
        with open(metadata_path, "r") as f:
            data = json.load(f)
        if "groups" not in data:
            raise ValueError("Invalid metadata format: missing 'groups' key")
        self.raw_groups = data["groups"]
        self.preference_pairs = select_preference_pairs(self.raw_groups, self.base_path, metric_name, metric_mode, min_gap,
                                                        metric_threshold, motion_threshold)
        if max_samples is not None:
            self.preference_pairs = self.preference_pairs[:max_samples]

    def __len__(self) -> int:
        return len(self.preference_pairs)

    def __getitem__(self, idx: int) -> Dict[str, Any]:
        pair = self.preference_pairs[idx]
        winner, loser = pair["winner"], pair["loser"]
        cond = _load(self.base_path / winner["condition_path"])            # the pair shares the winner's condition
        out = {"x_win": _load(self.base_path / winner["latent_path"]), "x_lose": _load(self.base_path / loser["latent_path"]),
               "prompt_emb": cond.get("encoder_hidden_states"), "prompt": pair["prompt"],
               "m_win": winner[self.metric_name], "m_lose": loser[self.metric_name]}
        if cond.get("image_embeds") is not None:
            out["image_emb"] = cond["image_embeds"]
        if cond.get("image_latent") is not None:
            out["image_latent"] = cond["image_latent"]
        return out


def collate_fn(batch: List[Dict[str, Any]]) -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    for key in ("x_win", "x_lose", "prompt_emb"):
        if key in batch[0]:
            out[key] = torch.stack([b[key] for b in batch])
    for key in ("image_emb", "image_latent"):
        if key in batch[0] and batch[0][key] is not None:
            out[key] = torch.stack([b[key] for b in batch])
    if "prompt" in batch[0]:
        out["prompt"] = [b["prompt"] for b in batch]
    for key in ("m_win", "m_lose"):
        if key in batch[0]:
            out[key] = torch.tensor([b[key] for b in batch])
    return out
