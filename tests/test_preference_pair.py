"""CPU tests of the preference-scoring driver mirror (videogpa_b200.train.preference_pair <-> train/01_preference_pair.py) and of its
hand-over to the dataset: the JSON it writes is what DPODataset's pair selection reads."""
import json
import os


class FakeProcessor:
    def __init__(self, scores):
        self.scores, self.calls = scores, []

    def process(self, video_path, thresholds, num_frames, save_visuals=False, out_dir=None):
        self.calls.append(os.path.basename(video_path))
        if "broken" in video_path:
            raise RuntimeError("cannot decode")
        cs, mn = self.scores[os.path.basename(video_path)]
        return {thresholds[0]: {"Consistency_Score": cs, "motion_norm": mn}}


def test_scoring_resume_and_formats(tmp_path):
    from videogpa_b200.train import preference_pair as pp
    vids = tmp_path / "v"
    vids.mkdir()
    for n in ("a0.mp4", "a1.mp4", "b0.mp4", "broken.mp4"):
        (vids / n).write_bytes(b"data")
    (vids / "empty.mp4").write_bytes(b"")
    groups = [
        {"group_id": "A", "text_prompt": "pa", "videos": [{"video_path": str(vids / "a0.mp4")}, {"video_path": str(vids / "a1.mp4")},
                                                           {"video_path": str(vids / "a0.mp4"), "note": "duplicate"}, {"note": "no path"}]},
        {"group_id": "B", "text_prompt": "pb", "videos": [{"video_path": str(vids / "b0.mp4")}, {"video_path": str(vids / "missing.mp4")},
                                                           {"video_path": str(vids / "empty.mp4")}, {"video_path": str(vids / "broken.mp4")}]},
        {"group_id": "C", "text_prompt": "pc", "videos": []},
    ]
    inp, out = tmp_path / "in.json", tmp_path / "out.json"
    inp.write_text(json.dumps({"groups": groups}))
    fake = FakeProcessor({"a0.mp4": (0.20, 0.5), "a1.mp4": (0.90, 0.6), "b0.mp4": (0.40, 0.7)})
    res = pp.process_video_scoring(str(inp), str(out), processor=fake)
    assert fake.calls == ["a0.mp4", "a1.mp4", "b0.mp4", "broken.mp4"]          # duplicate, missing, empty and path-less entries are not scored
    assert [g["group_id"] for g in res] == ["A", "B"]                            # the group without videos is dropped
    a, b = res
    assert a["videos"][0]["consistency_score"] == 0.20 and a["videos"][1]["motion_norm"] == 0.6
    assert "consistency_score" not in a["videos"][2] and a["videos"][3] == {"note": "no path"}
    assert b["videos"][0]["consistency_score"] == 0.40 and all("consistency_score" not in v for v in b["videos"][1:])
    assert json.loads(out.read_text()) == res and not os.path.exists(str(out) + ".tmp")
    # resume: nothing that already carries both scores is scored again; the plain-list input format works too
    inp.write_text(json.dumps(groups))
    fake2 = FakeProcessor({})
    res2 = pp.process_video_scoring(str(inp), str(out), processor=fake2)
    assert fake2.calls == ["broken.mp4"] and res2[0]["videos"][0]["consistency_score"] == 0.20
    # unusable inputs
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps({"items": []}))
    assert pp.process_video_scoring(str(bad), str(out), processor=fake2) is None
    assert pp.process_video_scoring(str(tmp_path / "nope.json"), str(out), processor=fake2) is None
    assert pp.extract_groups(3) is None and pp.extract_groups([1]) == [1]


def test_output_feeds_pair_selection(tmp_path):
    """The scored metadata is the input of videogpa_b200.dataset.select_preference_pairs (train/dataset.py:102-201): best vs worst by
    consistency_score (mode min), min_gap and motion threshold."""
    from videogpa_b200.dataset import select_preference_pairs
    from videogpa_b200.train import preference_pair as pp
    vids = tmp_path / "v"
    vids.mkdir()
    entries = []
    for n, lat in (("w.mp4", "w.pt"), ("l.mp4", "l.pt")):
        (vids / n).write_bytes(b"data")
        (tmp_path / lat).write_bytes(b"latent")
        entries.append({"video_path": str(vids / n), "latent_path": lat, "condition_path": "c.pt"})
    (tmp_path / "c.pt").write_bytes(b"cond")
    inp, out = tmp_path / "in.json", tmp_path / "out.json"
    inp.write_text(json.dumps([{"group_id": "G", "text_prompt": "p", "videos": entries}]))
    pp.process_video_scoring(str(inp), str(out), processor=FakeProcessor({"w.mp4": (0.10, 0.5), "l.mp4": (0.80, 0.5)}))
    groups = json.loads(out.read_text())
    pairs = select_preference_pairs(groups, tmp_path, metric_name="consistency_score", metric_mode="min", min_gap=0.05, motion_threshold=0.001)
    assert len(pairs) == 1 and pairs[0]["winner"]["latent_path"] == "w.pt" and pairs[0]["loser"]["latent_path"] == "l.pt"
    assert abs(pairs[0]["metric_gap"] - 0.70) < 1e-12 and pairs[0]["prompt"] == "p"
