"""CPU: the training loop around the DPO step (train/CogVideoX-5B/03_train.py:252-287 without Lightning) — stopping rule,
accumulation windows, equal-sized DDP shards, checkpoint retention. The step object is a stand-in: the loop is host logic."""
import os
import subprocess
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeStep:
    """Quacks like train_step.DPOSharedStep for `fit`: one scalar parameter, loss = w * x."""

    def __init__(self):
        self.w = torch.nn.Parameter(torch.ones(()))
        self.trainable = types.SimpleNamespace(parameters=lambda: [self.w])
        self.last_output = types.SimpleNamespace(reward_margin=torch.zeros(1))
        self.seen = []

    def configure_optimizers(self, lr):
        return torch.optim.SGD([self.w], lr=lr)

    def training_step(self, batch):
        self.seen.append(float(batch))
        return self.w * float(batch)

    def validation_step(self, batch):
        return {"val/loss": torch.tensor(float(batch))}


def _cfg(**kw):
    from videogpa_b200.train import cogvideox_5b as t
    cfg = dict(t.DEFAULT_CONFIG)
    cfg.update(learning_rate=0.0, warmup_steps=0, log_every_n_steps=1000, **kw)
    return cfg


def test_fit_runs_to_max_steps_and_ignores_max_epochs():
    from videogpa_b200.train import cogvideox_5b as t
    step = FakeStep()
    # 5 batches per epoch, accumulation 2 -> optimizer steps after batches 2, 4 and (flush) 5: 3 per epoch
    res = t.fit(step, [1.0, 2.0, 3.0, 4.0, 5.0], _cfg(max_steps=7, max_epochs=1, accumulate_grad_batches=2), val_loader=[3.0, 5.0], log=lambda *_: None)
    assert res["steps"] == 7                      # max_epochs = 1 would have stopped at 3 (Lightning: max_epochs = -1 with max_steps)
    assert res["epochs"] == 3 and res["val"] == [4.0, 4.0, 4.0]
    # epochs 1-2 complete (5 batches each); epoch 3 stops on its first window (batches 1, 2)
    assert step.seen == [1.0, 2.0, 3.0, 4.0, 5.0] * 2 + [1.0, 2.0]


def test_fit_flushes_partial_accumulation_at_epoch_end():
    from videogpa_b200.train import cogvideox_5b as t
    step = FakeStep()
    cfg = _cfg(max_steps=2, accumulate_grad_batches=4, gradient_clip_val=None)
    cfg["learning_rate"] = 1.0
    t.fit(step, [1.0, 2.0, 3.0, 4.0, 5.0], cfg, log=lambda *_: None)
    # step 1 (lr 1): window of batches 1-4, grad = (1+2+3+4)/4 = 2.5; step 2 (flush; cosine multiplier 0.5 half-way): batch 5 alone, 5/4
    assert abs(float(step.w.detach()) - (1.0 - 2.5 - 0.5 * 1.25)) < 1e-6


def test_padded_shards_are_equal_sized():
    from videogpa_b200.parallel import shard_padded
    for n, world in [(5, 2), (58, 8), (8, 8), (3, 8), (1, 4)]:
        shards = [shard_padded(range(n), r, world) for r in range(world)]
        assert len({len(s) for s in shards}) == 1 and len(shards[0]) == -(-n // world)
        assert set(sum(shards, [])) == set(range(n))                    # every item is still covered
    # same rule as torch's DistributedSampler(shuffle=False)
    from torch.utils.data.distributed import DistributedSampler
    for r in range(3):
        assert shard_padded(range(10), r, 3) == list(DistributedSampler(range(10), num_replicas=3, rank=r, shuffle=False))
    assert shard_padded([], 0, 2) == []


def test_topk_checkpoints_keep_lowest_val_loss(tmp_path):
    from videogpa_b200.train import cogvideox_5b as t
    saved = []

    def save(path):
        os.makedirs(path)
        saved.append(path)

    ck = t.TopKCheckpoints(tmp_path, save, top_k=2)
    ck(1000, float("nan")); ck(2000, 0.7); ck(3000, 0.5); ck(4000, 0.9)
    left = sorted(os.listdir(tmp_path))
    assert left == ["step=2000-val_loss=0.7000", "step=3000-val_loss=0.5000"] and len(saved) == 4


_WORKER = r'''
import os, sys, types, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VGPA_ROOT"])
sys.path.insert(0, os.path.join(os.environ["VGPA_ROOT"], "tests"))
from videogpa_b200.parallel import init_from_env, shard_padded
from videogpa_b200.train import cogvideox_5b as t
from test_host_train_loop import FakeStep, _cfg
rank, world, _ = init_from_env("gloo")
items = [float(i + 1) for i in range(5)]                      # 5 pairs over 2 ranks: 3 each after padding
mine = [items[i] for i in shard_padded(range(5), rank, world)]
assert len(mine) == 3
step = FakeStep()
cfg = _cfg(max_steps=5, accumulate_grad_batches=2, gradient_clip_val=None)
cfg["learning_rate"] = 0.5
res = t.fit(step, mine, cfg, log=lambda *_: None, rank=rank)     # hung before: ranks issued different numbers of all-reduces
assert res["steps"] == 5
ws = [torch.zeros(()) for _ in range(world)]
dist.all_gather(ws, step.w.detach())
assert torch.equal(ws[0], ws[1])                                # averaged gradients -> identical parameters on both ranks
dist.barrier(); dist.destroy_process_group()
print("loop ok", rank)
'''


def test_fit_world_size_2_gloo_uneven_dataset(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, VGPA_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", str(script)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("loop ok") == 2
