"""CPU tests of the scorer driver mirror (videogpa_b200.score <-> replicate_scorer.py): configuration from the SCORE_* environment,
task collection, chunking, per-video items (success and failure), resume, CSV / JSON report and summary."""
import json
import os

import pytest


class FakeProcessor:
    """Stands in for VideoProcessor.process: deterministic numbers from the file name, one failing video."""

    def __init__(self):
        self.calls = []

    def process(self, video_path, thresholds, num_frames, save_visuals=False):
        self.calls.append((video_path, tuple(thresholds), num_frames))
        name = os.path.basename(video_path)
        if "bad" in name:
            raise RuntimeError("decode failed")
        k = float(len(name))
        return {thresholds[0]: {"MSE": 0.01 * k, "Consistency_Score": 0.1 * k, "motion_norm": 0.5, "MVCS": 0.9, "Epipolar": 1.5}}


def _tree(root):
    for pid, names in {"p1": ["seed_42.mp4", "seed_456.mp4"], "p0": ["seed_42.mp4", "bad_seed_42.mp4"], "empty": []}.items():
        d = root / pid
        d.mkdir(parents=True)
        for n in names:
            (d / n).write_bytes(b"x")
    (root / "stray.mp4").write_bytes(b"x")                      # files at the top level are not prompt directories


def test_config_from_environment(monkeypatch):
    from videogpa_b200 import score
    for k in list(os.environ):
        if k.startswith("SCORE_"):
            monkeypatch.delenv(k)
    c = score.build_score_config()
    assert c["devices"] == [0] and c["base_dir"] == "output/replicate" and c["output_csv"] == "output/replicate/scores.csv"
    assert c["num_frames"] == 10 and c["conf_thres"] == 0 and c["backbone"] == "da3" and c["model_name"] == score.DEFAULT_DA3_MODEL
    assert c["descriptor_type"] == "lightglue" and c["resume"] is False and c["max_videos"] == 0 and c["ignore_seed"] is True
    monkeypatch.setenv("SCORE_BACKBONE", "VGGT"); monkeypatch.setenv("SCORE_DEVICES", "0, 2,3"); monkeypatch.setenv("SCORE_RESUME", "yes")
    c = score.build_score_config()
    assert c["backbone"] == "vggt" and c["model_name"] == score.DEFAULT_VGGT_MODEL and c["devices"] == [0, 2, 3] and c["resume"] is True


def test_collect_chunk_score_and_report(tmp_path, capsys):
    from videogpa_b200 import score
    _tree(tmp_path / "videos")
    cfg = dict(score.build_score_config(), base_dir=str(tmp_path / "videos"), output_csv=str(tmp_path / "out" / "scores.csv"),
               output_json=str(tmp_path / "out" / "scores.json"), num_frames=7, conf_thres=0)
    tasks = score.collect_all_video_tasks(cfg)
    assert [t["relative_path"] for t in tasks] == ["p0/bad_seed_42.mp4", "p0/seed_42.mp4", "p1/seed_42.mp4", "p1/seed_456.mp4"]
    assert [t["relative_path"] for t in score.collect_all_video_tasks(dict(cfg, seed_filter="456"))] == ["p1/seed_456.mp4"]
    assert len(score.collect_all_video_tasks(dict(cfg, max_videos=3))) == 3
    # replicate_scorer.py:238-244: ceil(n / workers) contiguous chunks, padded with empty ones
    assert [len(c) for c in score.chunk_tasks(tasks, 3)] == [2, 2, 0]
    assert [len(c) for c in score.chunk_tasks(tasks, 1)] == [4] and score.chunk_tasks([], 2) == [[], []]
    fake = FakeProcessor()
    df = score.main(processor_factory=lambda c, gpu: fake, config=cfg)
    assert fake.calls[0][1:] == ((0,), 7) and len(fake.calls) == 4
    payload = json.loads((tmp_path / "out" / "scores.json").read_text())
    items = payload["items"]
    assert [(i["prompt_id"], i["video_name"]) for i in items] == [("p0", "bad_seed_42.mp4"), ("p0", "seed_42.mp4"), ("p1", "seed_42.mp4"),
                                                                  ("p1", "seed_456.mp4")]
    bad, good = items[0], items[1]
    assert bad["error"] == "decode failed" and all(bad[m] is None for m in score.METRIC_COLS)
    assert good["mvcs"] == 0.9 and good["epipolar"] == 1.5 and good["psnr"] == 0.0 and good["motion_score"] == 0.5 and good["backbone"] == "da3"
    assert abs(good["consistency_score"] - 0.1 * len("seed_42.mp4")) < 1e-12
    s = payload["summary"]["overall"]
    assert s["video_count"] == 4 and abs(s["mvcs"] - 0.9) < 1e-12 and s["psnr"] == 0.0          # means skip the failed video's None
    csv_text = (tmp_path / "out" / "scores.csv").read_text()
    assert csv_text.splitlines()[0].startswith("prompt_id,video_name,video_path,relative_path,backbone") and len(csv_text.splitlines()) == 5
    assert "Overall Mean Metrics" in capsys.readouterr().out and len(df) == 4
    # resume: scored videos are skipped, the report keeps them
    fake2 = FakeProcessor()
    (tmp_path / "videos" / "p1" / "seed_7.mp4").write_bytes(b"x")
    score.main(processor_factory=lambda c, gpu: fake2, config=dict(cfg, resume=True))
    assert [os.path.basename(c[0]) for c in fake2.calls] == ["seed_7.mp4"]
    assert len(json.loads((tmp_path / "out" / "scores.json").read_text())["items"]) == 5


def test_missing_backbone_and_empty_dir(tmp_path, capsys):
    from videogpa_b200 import score
    cfg = dict(score.build_score_config(), base_dir=str(tmp_path / "none"))
    assert score.main(config=cfg) is None and "No videos found" in capsys.readouterr().out
    _tree(tmp_path / "videos")
    with pytest.raises(RuntimeError, match="backbone_fn"):
        score.main(config=dict(cfg, base_dir=str(tmp_path / "videos"), output_csv="", output_json=""))
