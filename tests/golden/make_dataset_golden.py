"""Generates tests/golden/dataset_pairs.json by running the REFERENCE's own train/dataset.py (DPODataset) in the build
container on a synthetic metadata file that exercises every selection rule. Run from the repo root:
    python tests/golden/make_dataset_golden.py
The fixture stores the metadata (so the test can rebuild the files), and for several parameter sets the selected
(group_id, winner generation_id, loser generation_id, metric_gap) tuples plus one loaded item's tensor checksums."""
import importlib.util
import json
import os
import sys
import tempfile

import torch

REF = "/root/reference/train/dataset.py"


def build_metadata():
    def vid(gid, k, score, motion, **kw):
        v = {"video_path": f"gen/{gid}_{k}.mp4", "generation_id": str(k), "consistency_score": score, "motion_norm": motion,
             "latent_path": f"lat/{gid}_{k}.pt", "condition_path": f"lat/{gid}_{k}_c.pt"}
        v.update(kw)
        return v
    groups = [
        {"group_id": "normal", "prompt": "a cat", "input_image_path": "i/cat.png", "videos": [vid("normal", 1, 3.5, 1.2), vid("normal", 2, 8.7, 1.1), vid("normal", 3, 5.0, 0.9)]},
        {"group_id": "text_prompt_key", "text_prompt": "a dog", "image_path": "i/dog.png", "videos": [vid("tp", 1, 0.9, 0.5), vid("tp", 2, 0.2, 0.5)]},
        {"group_id": "single", "prompt": "x", "videos": [vid("single", 1, 1.0, 1.0)]},
        {"group_id": "small_gap", "prompt": "x", "videos": [vid("sg", 1, 1.00, 1.0), vid("sg", 2, 1.05, 1.0)]},
        {"group_id": "static", "prompt": "x", "videos": [vid("st", 1, 1.0, 0.0001), vid("st", 2, 3.0, 1.0), vid("st", 3, 2.0, 0.0)]},
        {"group_id": "missing_metric", "prompt": "x", "videos": [vid("mm", 1, 1.0, 1.0), {"video_path": "gen/mm_2.mp4", "generation_id": "2", "motion_norm": 1.0, "latent_path": "lat/mm_2.pt", "condition_path": "lat/mm_2_c.pt"}, vid("mm", 3, 4.0, 1.0)]},
        {"group_id": "missing_file", "prompt": "x", "videos": [vid("mf", 1, 1.0, 1.0), vid("mf", 2, 9.0, 1.0, _no_file=True), vid("mf", 3, 2.0, 1.0)]},
        {"group_id": "ties", "prompt": "x", "videos": [vid("ti", 1, 2.0, 1.0), vid("ti", 2, 1.0, 1.0), vid("ti", 3, 1.0, 1.0), vid("ti", 4, 2.0, 1.0)]},
        {"group_id": "threshold", "prompt": "x", "videos": [vid("th", 1, 6.0, 1.0), vid("th", 2, 9.0, 1.0)]},
        {"group_id": "no_paths", "prompt": "x", "videos": [{"video_path": "a", "consistency_score": 1.0, "motion_norm": 1.0}, {"video_path": "b", "consistency_score": 5.0, "motion_norm": 1.0}]},
        {"prompt": "no id", "videos": [vid("noid", 1, 0.5, 2.0), vid("noid", 2, 7.5, 2.0)]},
    ]
    return {"groups": groups}


def materialise(meta, root):
    """Write the latent / condition files the metadata points to (tiny tensors, seeded by the path)."""
    for g in meta["groups"]:
        for v in g.get("videos", []):
            if "latent_path" not in v or v.get("_no_file"):
                continue
            for key, kind in (("latent_path", "lat"), ("condition_path", "cond")):
                p = os.path.join(root, v[key])
                os.makedirs(os.path.dirname(p), exist_ok=True)
                gen = torch.Generator().manual_seed(sum(map(ord, v[key])))
                if kind == "lat":
                    torch.save(torch.randn(4, 3, 6, 8, generator=gen), p)
                else:
                    torch.save({"encoder_hidden_states": torch.randn(5, 16, generator=gen), "image_embeds": torch.randn(3, 6, 8, generator=gen)}, p)


PARAM_SETS = {
    "default": {},
    "max_mode": {"metric_mode": "max"},
    "tight_gap": {"min_gap": 1.5},
    "winner_threshold": {"metric_threshold": 3.0},
    "motion_metric": {"metric_name": "motion_norm", "metric_mode": "max", "min_gap": 0.05},
    "max_samples": {"max_samples": 2},
}


def main():
    spec = importlib.util.spec_from_file_location("ref_dataset", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    meta = build_metadata()
    out = {"metadata": meta, "selections": {}, "item0": {}}
    with tempfile.TemporaryDirectory() as root:
        materialise(meta, root)
        mp = os.path.join(root, "meta_data.json")
        json.dump(meta, open(mp, "w"))
        for name, kw in PARAM_SETS.items():
            ds = ref.DPODataset(root, mp, **kw)
            out["selections"][name] = [[p["group_id"], p["winner"]["generation_id"], p["loser"]["generation_id"], p["metric_gap"]] for p in ds.preference_pairs]
        ds = ref.DPODataset(root, mp)
        item = ds[0]
        out["item0"] = {k: (float(v.double().sum()) if torch.is_tensor(v) else v) for k, v in item.items()}
        batch = ref.collate_fn([ds[0], ds[1]])
        out["collate_keys"] = sorted(batch.keys())
        out["collate_shapes"] = {k: list(v.shape) for k, v in batch.items() if torch.is_tensor(v)}
    json.dump(out, open(os.path.join(os.path.dirname(__file__), "dataset_pairs.json"), "w"), indent=1)
    print({k: len(v) for k, v in out["selections"].items()})


if __name__ == "__main__":
    main()
