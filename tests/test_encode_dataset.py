"""CPU tests of the 02_encode.py mirror (videogpa_b200.train.encode_dataset) with stand-in encoders: file naming, relative paths, the
unscaled / scaled latent rule, skipped groups and videos, and the hand-over chain scorer JSON -> encoder JSON -> DPODataset."""
import json
from types import SimpleNamespace

import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")


class FakeVAE:
    """vae.encode(x).latent_dist.sample(): 4x temporal / 8x spatial average pooling of the first channel, 16 channels."""
    device, dtype = torch.device("cpu"), torch.float32
    config = SimpleNamespace(scaling_factor=0.7)

    def encode(self, x):                                          # x [1, 3, T, H, W] in [0, 1]
        assert x.dim() == 5 and x.shape[1] == 3 and 0.0 <= float(x.min()) and float(x.max()) <= 1.0
        T = (x.shape[2] - 1) // 4 + 1
        lat = torch.nn.functional.adaptive_avg_pool3d(x[:, :1], (T, x.shape[3] // 8, x.shape[4] // 8)).repeat(1, 16, 1, 1, 1)
        return SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda generator=None: lat))


class FakeT5:
    device = torch.device("cpu")

    def __call__(self, ids):
        return (ids.float().unsqueeze(-1).repeat(1, 1, 8),)      # [1, 226, 8]


def _clip(path, n=9, h=64, w=96):
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"mp4v"), 8.0, (w, h))
    for k in range(n):
        wr.write(np.full((h, w, 3), 20 + 10 * k, dtype=np.uint8))
    wr.release()


def test_encode_groups_and_chain_to_dataset(tmp_path):
    from videogpa_b200.dataset import select_preference_pairs
    from videogpa_b200.train import encode_dataset as E
    base = tmp_path / "root"
    (base / "videos").mkdir(parents=True)
    for n in ("a.mp4", "b.mp4"):
        _clip(base / "videos" / n)
    groups = [
        {"group_id": "G1", "text_prompt": "a street", "extra": 1,
         "videos": [{"video_path": "videos/a.mp4", "consistency_score": 0.1, "motion_norm": 0.5},
                    {"video_path": "videos/b.mp4", "consistency_score": 0.9, "motion_norm": 0.5},
                    {"video_path": "videos/missing.mp4", "consistency_score": 0.5, "motion_norm": 0.5}, {"note": "no path"}]},
        {"group_id": "G2", "text_prompt": "", "videos": [{"video_path": "videos/a.mp4"}]},          # no prompt: skipped
        {"group_id": "G3", "text_prompt": "only broken", "videos": [{"video_path": "videos/missing.mp4"}]},   # nothing encoded: dropped
    ]
    inp, out = base / "meta_temp.json", base / "meta_data.json"
    inp.write_text(json.dumps({"t2v_groups": groups}))
    tok = lambda prompt: torch.arange(226).unsqueeze(0) + len(prompt)
    res = E.process_t2v_encoding(str(inp), str(out), str(base), FakeVAE(), FakeT5(), tok, num_frames=49)
    assert json.loads(out.read_text()) == res and [g["group_id"] for g in res["groups"]] == ["G1"]
    g1 = res["groups"][0]
    assert set(g1) == {"group_id", "text_prompt", "videos"} and len(g1["videos"]) == 2
    v = g1["videos"][0]
    assert v["condition_path"] == "t2v_latent/cond_G1.pt" and v["latent_path"] == "t2v_latent/latent_G1_a.pt" and v["consistency_score"] == 0.1
    cond = torch.load(str(base / v["condition_path"]))
    assert set(cond) == {"encoder_hidden_states"} and tuple(cond["encoder_hidden_states"].shape) == (226, 8)
    assert float(cond["encoder_hidden_states"][0, 0]) == len("a street")
    lat = torch.load(str(base / v["latent_path"]))
    assert tuple(lat.shape) == (16, 3, 8, 12)                     # 9 frames (all of a short clip) -> 3 latent frames, 64x96 -> 8x12
    # the 1.5 rule: latent * scaling_factor
    res15 = E.process_t2v_encoding(str(inp), str(base / "m15.json"), str(base), FakeVAE(), FakeT5(), tok, latent_root=str(base / "lat15"), scale_latents=True)
    lat15 = torch.load(str(base / res15["groups"][0]["videos"][0]["latent_path"]))
    assert torch.allclose(lat15, lat * 0.7)
    # plain-list input, and what the dataset's pair selection makes of the output
    inp.write_text(json.dumps(groups))
    assert E.process_t2v_encoding(str(inp), str(out), str(base), FakeVAE(), FakeT5(), tok)["groups"][0]["group_id"] == "G1"
    pairs = select_preference_pairs(json.loads(out.read_text())["groups"], base, metric_name="consistency_score", metric_mode="min", min_gap=0.05)
    assert len(pairs) == 1 and pairs[0]["winner"]["latent_path"].endswith("latent_G1_a.pt") and pairs[0]["loser"]["latent_path"].endswith("latent_G1_b.pt")
    assert E.process_t2v_encoding(str(base / "none.json"), str(out), str(base), FakeVAE(), FakeT5(), tok) is None
    assert E.extract_groups({"groups": [1]}) == [1] and E.extract_groups(5) == []
    # the I2V variant: image_embeds in the condition file, files under <latent_root>/processed, whole group kept, groups without image skipped
    from PIL import Image
    Image.fromarray(np.full((40, 60, 3), 128, dtype=np.uint8)).save(base / "first.png")
    gi = [dict(groups[0], image_path="first.png"), groups[0]]
    inp.write_text(json.dumps({"groups": gi}))
    ri = E.process_t2v_encoding(str(inp), str(base / "meta_i2v.json"), str(base), FakeVAE(), FakeT5(), tok, latent_root=str(base / "i2v_latent"),
                                image_condition=True, sub_folder="processed")
    assert len(ri["groups"]) == 1 and ri["groups"][0]["extra"] == 1 and ri["groups"][0]["image_path"] == "first.png"
    vi = ri["groups"][0]["videos"][0]
    assert vi["condition_path"] == "i2v_latent/processed/cond_G1.pt"
    ci = torch.load(str(base / vi["condition_path"]))
    assert tuple(ci["image_embeds"].shape) == (3, 40, 60) and abs(float(ci["image_embeds"].mean()) - 128 / 255) < 1e-6
