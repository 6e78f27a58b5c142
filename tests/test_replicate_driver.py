"""The replicate.py mirror (videogpa_b200.replicate): CPU tests of the configuration / hash / first-frame / job-planning rules, and a
GPU end-to-end run in synthetic mode (random 1-block I2V model) with a real PEFT adapter directory whose strength changes per work item."""
import json
import os

import pytest


def _dataset(root, keys):
    for k in keys:
        d = root / k
        d.mkdir(parents=True)
        import numpy as np
        from PIL import Image
        rng = np.random.default_rng(len(k))
        Image.fromarray(rng.integers(0, 255, (90, 120, 3), dtype=np.uint8)).save(d / "frame_00001.png")


def test_config_hash_and_job_plan(tmp_path, monkeypatch):
    from videogpa_b200 import replicate as R
    for k in ("RUN_MODE", "RUN_LORA_PATH", "RUN_OUTPUT_DIR", "PROMPT_JSON", "DL3DV_BASE_DIR", "RUN_DEVICES", "RUN_NUM_PROMPTS", "RUN_SEEDS"):
        monkeypatch.delenv(k, raising=False)
    c = R.build_config(here="/repo")
    assert c["mode"] == "dpo" and c["devices"] == [0] and c["num_prompts"] == 100 and c["seeds_per_prompt"] == [456] and c["weight_list"] == [1.0]
    assert c["lora_path"] == "/repo/checkpoints/VideoGPA-I2V-lora" and c["base_model"] == "THUDM/CogVideoX-5B-I2V"
    assert c["num_inference_steps"] == 50 and c["guidance_scale"] == 6.0 and c["fps"] == 8
    monkeypatch.setenv("RUN_SEEDS", "1, 2"); monkeypatch.setenv("RUN_MODE", "base"); monkeypatch.setenv("RUN_DEVICES", "0,1,2,3")
    c = R.build_config(here="/repo")
    assert c["seeds_per_prompt"] == [1, 2] and c["mode"] == "base" and c["devices"] == [0, 1, 2, 3] and R.weights_for(c) == [0.0]
    # replicate.py:46-63
    assert R.extract_pure_hash_from_json_key("1K/abc123/images_8") == "abc123"
    assert R.extract_pure_hash_from_json_key(" scene/42 ") == "scene_42" and R.extract_pure_hash_from_json_key("plain") == "plain"
    with pytest.raises(RuntimeError):
        R.extract_pure_hash_from_json_key("a//b")
    assert R.video_filename("dpo", 456, 1.0) == "seed_456_dpo_w1.0.mp4" and R.video_filename("base", 7, 0.0) == "seed_7_original.mp4"
    # job plan: item -> weight -> seed; empty prompts and missing frame folders are skipped; the shard is items[r::world]
    data = tmp_path / "dl3dv"
    _dataset(data, ["1K/h0/images_8", "1K/h1/images_8", "1K/h2/images_8"])
    (data / "1K" / "h3" / "images_8").mkdir(parents=True)                      # folder without frame_00001.png
    caps = {"1K/h0/images_8": " a street ", "1K/h1/images_8": "  ", "1K/h2/images_8": "a park", "1K/h3/images_8": "no frame", "1K/h4/images_8": "no dir"}
    pj = tmp_path / "caps.json"
    pj.write_text(json.dumps(caps))
    cfg = dict(R.build_config(here=str(tmp_path)), mode="dpo", weight_list=[0.5, 1.0], seeds_per_prompt=[3, 4], prompt_json=str(pj),
               dl3dv_base_dir=str(data), output_dir=str(tmp_path / "out"), num_prompts=4)
    items = R.select_items(cfg)
    assert [k for k, _ in items] == list(caps)[:4]
    msgs = []
    jobs = R.plan_jobs(items, cfg, log=msgs.append)
    assert [(j["pure_hash"], j["lora_weight"], j["seed"]) for j in jobs] == [("h0", 0.5, 3), ("h0", 0.5, 4), ("h0", 1.0, 3), ("h0", 1.0, 4),
                                                                             ("h2", 0.5, 3), ("h2", 0.5, 4), ("h2", 1.0, 3), ("h2", 1.0, 4)]
    assert jobs[0]["prompt"] == "a street" and str(jobs[2]["path"]).endswith(os.path.join("out", "h0", "seed_3_dpo_w1.0.mp4"))
    assert len(msgs) == 2 and "empty prompt" in msgs[0] and "frame_00001.png" in msgs[1]
    assert [k for k, _ in items[1::2]] == ["1K/h1/images_8", "1K/h3/images_8"]    # rank 1 of 2
    assert R.main(config=dict(cfg, dl3dv_base_dir=str(tmp_path / "nope")), argv=[]) == 0


@pytest.mark.gpu
def test_replicate_worker_synthetic_with_lora_weights(lib, tmp_path, capsys):
    import torch
    from safetensors.torch import save_file
    from videogpa_b200 import replicate as R
    data = tmp_path / "dl3dv"
    _dataset(data, ["1K/h0/images_8"])
    pj = tmp_path / "caps.json"
    pj.write_text(json.dumps({"1K/h0/images_8": "a slow pan over a street", "1K/h9/images_8": "missing folder"}))
    # a PEFT adapter directory for the 1-block synthetic I2V model (rank 8)
    lora = tmp_path / "lora"
    lora.mkdir()
    g = torch.Generator().manual_seed(0)
    tensors = {}
    for mod in ("to_q", "to_k", "to_v", "to_out.0"):
        base = f"base_model.model.transformer_blocks.0.attn1.{mod}"
        tensors[base + ".lora_A.weight"] = torch.randn(8, 3072, generator=g) * 0.05
        tensors[base + ".lora_B.weight"] = torch.randn(3072, 8, generator=g) * 0.05
    save_file(tensors, str(lora / "adapter_model.safetensors"))
    (lora / "adapter_config.json").write_text(json.dumps(dict(peft_type="LORA", r=8, lora_alpha=16.0, target_modules=["to_q", "to_k", "to_v", "to_out.0"],
                                                              use_dora=False, use_rslora=False, fan_in_fan_out=False, bias="none")))
    cfg = dict(R.build_config(here=str(tmp_path)), mode="dpo", weight_list=[0.0, 1.0], seeds_per_prompt=[7], prompt_json=str(pj), dl3dv_base_dir=str(data),
               output_dir=str(tmp_path / "out"), lora_path=str(lora), num_prompts=5, num_inference_steps=2, devices=[0])
    n = R.main(argv=["--synthetic", "1"], config=cfg)
    txt = capsys.readouterr().out
    assert n == 2 and "skipping entry" in txt and "generation failed" not in txt, txt
    out = tmp_path / "out" / "h0"
    assert sorted(p.name for p in out.iterdir()) == ["seed_7_dpo_w0.0.mp4", "seed_7_dpo_w1.0.mp4"]
    a, b = (out / "seed_7_dpo_w0.0.mp4").read_bytes(), (out / "seed_7_dpo_w1.0.mp4").read_bytes()
    assert len(a) > 1000 and a != b                                  # same seed, different adapter strength: different video
    assert R.main(argv=["--synthetic", "1"], config=cfg) == 0 and "already exists" in capsys.readouterr().out     # resume
