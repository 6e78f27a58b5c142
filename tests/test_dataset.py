"""CPU: the DPO dataset mirror (videogpa_b200/dataset.py) against a fixture produced by the REFERENCE's own
train/dataset.py (tests/golden/make_dataset_golden.py -> tests/golden/dataset_pairs.json): pair selection under every rule
(too few videos, missing metric / paths / files, static videos, small gap, winner threshold, ties, max mode, max_samples),
item loading and collation."""
import json
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture(scope="module")
def fixture(tmp_path_factory):
    import make_dataset_golden as mk
    gold = json.load(open(os.path.join(HERE, "golden", "dataset_pairs.json")))
    root = tmp_path_factory.mktemp("dpo")
    mk.materialise(gold["metadata"], str(root))
    mp = root / "meta_data.json"
    mp.write_text(json.dumps(gold["metadata"]))
    return gold, str(root), str(mp), mk.PARAM_SETS


def test_pair_selection_matches_reference(fixture):
    from videogpa_b200.dataset import DPODataset
    gold, root, mp, param_sets = fixture
    assert gold["metadata"] == __import__("make_dataset_golden").build_metadata()       # fixture and generator agree
    for name, kw in param_sets.items():
        ds = DPODataset(root, mp, **kw)
        got = [[p["group_id"], p["winner"]["generation_id"], p["loser"]["generation_id"], p["metric_gap"]] for p in ds.preference_pairs]
        assert got == gold["selections"][name], name
        assert len(ds) == len(gold["selections"][name])


def test_item_and_collate_match_reference(fixture):
    from videogpa_b200.dataset import DPODataset, collate_fn
    gold, root, mp, _ = fixture
    ds = DPODataset(root, mp)
    item = ds[0]
    for k, want in gold["item0"].items():
        got = float(item[k].double().sum()) if torch.is_tensor(item[k]) else item[k]
        assert got == want, k
    batch = collate_fn([ds[0], ds[1]])
    assert sorted(batch.keys()) == gold["collate_keys"]
    assert {k: list(v.shape) for k, v in batch.items() if torch.is_tensor(v)} == gold["collate_shapes"]
    assert batch["m_win"].dtype == torch.float32 and isinstance(batch["prompt"], list)


def test_invalid_metadata_raises(tmp_path):
    from videogpa_b200.dataset import DPODataset
    p = tmp_path / "m.json"
    p.write_text(json.dumps({"not_groups": []}))
    with pytest.raises(ValueError):
        DPODataset(str(tmp_path), str(p))
