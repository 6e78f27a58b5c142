"""CPU: host-side logic that needs no GPU — scheduler tables, RoPE tables, adapter parsing, sharding rules
and the world_size-2 exchange paths over gloo."""
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import dit_torch as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scheduler_tables_match_oracle():
    from videogpa_b200.schedulers import CogVideoXDDIMScheduler, CogVideoXDPMScheduler
    s = CogVideoXDDIMScheduler()
    ts = s.set_timesteps(50)
    assert ts.tolist() == list(range(999, 0, -20)) == O.trailing_timesteps(50).tolist()
    ac = O.cogvideox_alphas_cumprod()
    assert np.array_equal(s.alphas_cumprod, ac)
    assert ac[-1] == 0.0 and 0.99 < ac[0] < 1.0 and np.all(np.diff(ac) < 0)      # zero terminal SNR, monotone
    # DDIM coefficients reproduce the closed form (App. A.4), including the final step to alpha_prev = 1
    for t in (999, 499, 19):
        k = s.coefficients(t)
        a_t = ac[t]; a_p = ac[t - 20] if t - 20 >= 0 else 1.0
        a = math.sqrt((1 - a_p) / (1 - a_t))
        assert abs(k["c_sample"] - a) < 1e-15 and abs(k["c_x0"] - (math.sqrt(a_p) - math.sqrt(a_t) * a)) < 1e-15
    k = s.coefficients(19)
    assert k["c_sample"] == 0.0 and abs(k["c_x0"] - 1.0) < 1e-15                 # last step returns x0
    # at t = 999 alpha_bar = 0: x0 = -v
    k = s.coefficients(999)
    assert k["sqrt_alpha_t"] == 0.0 and k["sqrt_beta_t"] == 1.0
    d = CogVideoXDPMScheduler(); d.set_timesteps(50)
    m1, m2, mn, r = O.dpm_coefficients(ac, 499, 479, 519)
    k = d.coefficients(499, 519)
    assert abs(k["c_sample"] - m1) < 1e-14 and abs(k["c_noise"] - mn) < 1e-14
    assert abs(k["c_x0"] + m2 * (1 + 1 / (2 * r))) < 1e-14 and abs(k["c_x0_old"] - m2 / (2 * r)) < 1e-14
    k = d.coefficients(19, 39)                                                   # last step: x_prev = x0, no noise
    assert k["c_sample"] == 0.0 and k["c_x0"] == 1.0 and k["c_noise"] == 0.0 and k["c_x0_old"] == 0.0
    # add_noise / get_velocity (03_train.py:129-130,154-155)
    gsd = torch.Generator().manual_seed(0)
    x, n = torch.randn(2, 3, 4, generator=gsd), torch.randn(2, 3, 4, generator=gsd)
    t = torch.tensor([10, 900])
    assert torch.allclose(s.add_noise(x, n, t), O.add_noise(ac, x, n, t.numpy()))
    assert torch.allclose(s.get_velocity(x, n, t), O.get_velocity(ac, x, n, t.numpy()))


def test_scheduler_from_checkpoint_config(tmp_path):
    """generate/CogVideoX-5B.py:18 (`from_config(pipe.scheduler.config, timestep_spacing="trailing")`) and
    train/CogVideoX-5B/03_train.py:113 (`from_pretrained(model_path, subfolder="scheduler")`): the checkpoint's own betas and
    snr_shift_scale decide the noise levels (CogVideoX-2B ships snr_shift_scale 3.0)."""
    from videogpa_b200.schedulers import CogVideoXDDIMScheduler, CogVideoXDPMScheduler
    cfg = {"_class_name": "CogVideoXDDIMScheduler", "_diffusers_version": "0.30.0", "beta_start": 0.00085, "beta_end": 0.012,
           "beta_schedule": "scaled_linear", "clip_sample": False, "num_train_timesteps": 1000, "prediction_type": "v_prediction",
           "rescale_betas_zero_snr": True, "set_alpha_to_one": True, "snr_shift_scale": 3.0, "steps_offset": 0,
           "timestep_spacing": "linspace", "trained_betas": None}
    (tmp_path / "scheduler").mkdir()
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(cfg))
    d = CogVideoXDPMScheduler.from_pretrained(str(tmp_path), timestep_spacing="trailing")
    assert np.array_equal(d.alphas_cumprod, O.cogvideox_alphas_cumprod(snr_shift_scale=3.0))
    assert not np.array_equal(d.alphas_cumprod, CogVideoXDPMScheduler().alphas_cumprod)
    s = CogVideoXDDIMScheduler.from_config(dict(cfg, snr_shift_scale=1.0, timestep_spacing="trailing"))
    assert np.array_equal(s.alphas_cumprod, CogVideoXDDIMScheduler().alphas_cumprod)
    with pytest.raises(RuntimeError):
        CogVideoXDDIMScheduler.from_config(cfg)                                  # linspace spacing is not implemented: loud, not silent
    with pytest.raises(RuntimeError):
        CogVideoXDDIMScheduler.from_config(dict(cfg, timestep_spacing="trailing", prediction_type="epsilon"))
    with pytest.raises(RuntimeError):
        CogVideoXDPMScheduler.from_pretrained(str(tmp_path / "missing"))


def test_i2v_checkpoint_scheduler_selection(tmp_path):
    """generate/CogVideoX-5B-I2V.py:16-19 and replicate.py:158-162 keep the scheduler the checkpoint ships (no swap to DPM): the class
    comes from scheduler_config.json's `_class_name`."""
    from videogpa_b200.generate.cogvideox_5b_i2v import checkpoint_scheduler
    from videogpa_b200.schedulers import CogVideoXDDIMScheduler, CogVideoXDPMScheduler
    (tmp_path / "scheduler").mkdir()
    cfg = {"_class_name": "CogVideoXDDIMScheduler", "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear",
           "num_train_timesteps": 1000, "prediction_type": "v_prediction", "rescale_betas_zero_snr": True, "snr_shift_scale": 1.0,
           "timestep_spacing": "trailing"}
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(cfg))
    assert type(checkpoint_scheduler(str(tmp_path))) is CogVideoXDDIMScheduler
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(dict(cfg, _class_name="CogVideoXDPMScheduler")))
    assert type(checkpoint_scheduler(str(tmp_path))) is CogVideoXDPMScheduler
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(dict(cfg, _class_name="EulerDiscreteScheduler")))
    with pytest.raises(RuntimeError, match="not implemented"):
        checkpoint_scheduler(str(tmp_path))
    with pytest.raises(RuntimeError):
        checkpoint_scheduler(str(tmp_path / "missing"))


def test_rope_table_matches_oracle():
    from videogpa_b200.rope import get_3d_rotary_pos_embed
    cfg = O.DiTConfig()
    cos, sin = get_3d_rotary_pos_embed(64, 30, 45, 13)
    rc, rs = O.rope_3d(cfg, 13, 60, 90)
    assert cos.shape == (17550, 64) and torch.equal(cos, rc) and torch.equal(sin, rs)
    assert torch.all(cos[0] == 1) and torch.all(sin[0] == 0)                    # position (0,0,0)
    # t/h/w split 16/24/24, interleaved pairs share an angle
    assert torch.equal(cos[:, 0::2], cos[:, 1::2])
    assert torch.equal(cos[1, :16], cos[0, :16]) and not torch.equal(cos[1, 40:], cos[0, 40:])   # token 1 moves along w only


def test_transformer_config_and_flops():
    from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig
    c = TransformerConfig.cogvideox_5b()
    assert c.inner_dim == 3072 and c.num_layers == 42
    assert TransformerConfig.cogvideox_5b_i2v().in_channels == 32
    f = CogVideoXTransformer3D.flops_per_sample(type("M", (), {"config": c})(), 226, 17550)
    assert abs(f - 3.322e14) / 3.322e14 < 0.01                                   # 3.322e14 FLOP per sample-forward


def test_lora_adapter_reader(tmp_path):
    from safetensors.torch import save_file
    from videogpa_b200.lora import read_adapter
    t = {}
    for layer in range(2):
        for mod in ("to_q", "to_k", "to_v", "to_out.0"):
            base = f"base_model.model.transformer_blocks.{layer}.attn1.{mod}"
            t[base + ".lora_A.weight"] = torch.randn(64, 128)
            t[base + ".lora_B.weight"] = torch.randn(128, 64)
    t["base_model.model.something_else.weight"] = torch.zeros(1)
    save_file(t, str(tmp_path / "adapter_model.safetensors"))
    ref_cfg = json.load(open("/root/reference/checkpoints/VideoGPA-T2V-lora/adapter_config.json")) \
        if os.path.exists("/root/reference/checkpoints/VideoGPA-T2V-lora/adapter_config.json") else \
        dict(peft_type="LORA", r=64, lora_alpha=128.0, use_dora=False, use_rslora=False, fan_in_fan_out=False)
    (tmp_path / "adapter_config.json").write_text(json.dumps(ref_cfg))
    cfg, pairs = read_adapter(str(tmp_path))
    assert cfg["r"] == 64 and cfg["lora_alpha"] == 128.0 and len(pairs) == 8
    assert pairs[(1, "to_out.0")][0].shape == (64, 128) and pairs[(1, "to_out.0")][1].shape == (128, 64)
    bad = dict(ref_cfg, use_dora=True)
    (tmp_path / "adapter_config.json").write_text(json.dumps(bad))
    with pytest.raises(RuntimeError):
        read_adapter(str(tmp_path))
    with pytest.raises(RuntimeError):
        read_adapter(str(tmp_path / "nope"))


def test_sharding_rules():
    from videogpa_b200.parallel import shard_contiguous, shard_round_robin
    items = list(range(11))
    assert [shard_round_robin(items, r, 4) for r in range(4)] == [[0, 4, 8], [1, 5, 9], [2, 6, 10], [3, 7]]     # replicate.py:120
    chunks = [shard_contiguous(items, r, 4) for r in range(4)]
    assert sum(chunks, []) == items and [len(c) for c in chunks] == [3, 3, 3, 2]                                  # replicate_scorer.py:244-250
    assert shard_contiguous([], 0, 2) == [] and shard_round_robin([1], 1, 2) == []


def test_peer_pair_group_refuses_several_pairs():
    """The fused peer-memory exchange is validated for one CFG pair only (it hung with two pairs on 4 GPUs,
    profiles/r02_multi_gpu.md): more ranks must get an error, not a hang."""
    from videogpa_b200.parallel import CfgPairPeerGroup
    with pytest.raises(RuntimeError, match="one CFG pair"):
        CfgPairPeerGroup(0, 4)
    with pytest.raises(RuntimeError):
        CfgPairPeerGroup(1, 8)


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: CPU tensors are rejected by the host wrappers."""
    from videogpa_b200 import dense
    from videogpa_b200.geometry import batch_reproject
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        dense.linear(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):
        batch_reproject(np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32), np.zeros((1, 3, 3), np.float32),
                        np.zeros((1, 3, 4), np.float32), 4, 4)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VGPA_ROOT"])
from videogpa_b200.parallel import CfgPairGroup, average_gradients, gather_frames, gather_scores, init_from_env, shard_round_robin
rank, world, _ = init_from_env("gloo")
assert world == 2
grp = CfgPairGroup(rank, world)
assert grp.branch == rank and grp.pair == 0
pred = torch.full((1, 2, 3), float(rank + 1))
u, c = grp.exchange(pred)
assert torch.all(u == 1) and torch.all(c == 2)           # rank 0 = uncond, rank 1 = cond, identical on both ranks
# both ranks apply the same CFG + update -> identical latents
v = u + 6.0 * (c - u)
chk = [torch.zeros_like(v) for _ in range(2)]
dist.all_gather(chk, v)
assert torch.equal(chk[0], chk[1])
frames = torch.full((rank + 1, 2, 2, 3), rank, dtype=torch.uint8)
got = gather_frames(frames, rank, world)
if rank == 0:
    assert [g.shape[0] for g in got] == [1, 2] and int(got[1].max()) == 1
else:
    assert got is None
s = gather_scores(torch.arange(rank + 2, dtype=torch.float64) + 10 * rank, rank, world)
assert s.tolist() == [0.0, 1.0, 10.0, 11.0, 12.0]
assert shard_round_robin(range(5), rank, world) == ([0, 2, 4] if rank == 0 else [1, 3])
# DDP exchange of the training step: gradients averaged in place, a missing gradient counts as zero, buckets split by size
ps = [torch.zeros(3, 4, requires_grad=True), torch.zeros(5, requires_grad=True), torch.zeros(2, 2, requires_grad=True)]
ps[0].grad = torch.full((3, 4), float(rank + 1))
ps[1].grad = torch.arange(5.0) * (rank + 1)
if rank == 0:
    ps[2].grad = torch.ones(2, 2)
assert average_gradients(ps, bucket_bytes=64) == 2          # 12 + 5 floats exceed 64 bytes -> [p0], [p1, p2]
assert torch.all(ps[0].grad == 1.5) and torch.equal(ps[1].grad, torch.arange(5.0) * 1.5) and torch.all(ps[2].grad == 0.5)
# the overlapped reducer: same averages as average_gradients, launched from autograd hooks, silent while not armed
from videogpa_b200.parallel import BucketedGradReducer
qs = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
red = BucketedGradReducer(qs, bucket_bytes=64)
assert [sorted(b) for b in red.buckets] == [[1, 2], [0]]       # reverse (backward) order, 64-byte buckets
red.armed = False
((qs[0].sum() + qs[1].sum()) * float(rank + 1)).backward()     # accumulation micro-batch: no collective
assert not red._inflight
red.armed = True
((qs[0].sum() + qs[1].sum()) * float(rank + 1)).backward()     # q2 gets no gradient at all
st = red.finish()
assert st["buckets"] == 2 and st["bytes"] == (12 + 5 + 4) * 4
assert torch.all(qs[0].grad == 3.0) and torch.all(qs[1].grad == 3.0) and torch.all(qs[2].grad == 0.0)   # mean over ranks of 2 x (rank + 1)
red.remove()
dist.barrier(); dist.destroy_process_group()
print("worker ok", rank)
'''


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, VGPA_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("worker ok") == 2


def test_frame_and_pair_index_rules_match_oracle():
    """a-17: the bit-exact index rules of the host mirror against the oracle restatement and the SURVEY KAT."""
    from oracle import scorer_np as o
    from videogpa_b200.metrics import consecutive_pairs, sample_frame_indices
    assert sample_frame_indices(49, 10).tolist() == [0, 5, 10, 16, 21, 26, 32, 37, 42, 48]
    for total, n in [(49, 10), (6, 10), (1, 10), (81, 10), (120, 8), (10, 10)]:
        assert sample_frame_indices(total, n).tolist() == o.uniform_frame_indices(total, n).tolist()
    assert consecutive_pairs(4) == list(o.consecutive_pairs(4)) == [(0, 1), (1, 2), (2, 3)]
    assert consecutive_pairs(1) == []


def test_generate_cli_surface_and_task_parsing(tmp_path):
    """The generate CLI keeps the reference's flags/defaults (generate/CogVideoX-5B.py:86-99), prompt-JSON forms and
    output layout (host logic only; no GPU)."""
    import json
    from videogpa_b200.generate import cogvideox_5b as g
    p = g.build_parser()
    a = p.parse_args(["--prompt_json", "x.json", "--output_dir", "out"])
    assert (a.base_model, a.lora_path, a.gpu_id, a.seed, a.num_prompts, a.num_inference_steps, a.guidance_scale, a.fps) == \
        ("THUDM/CogVideoX-5B", None, 0, 42, None, 50, 6.0, 8)
    import pytest
    with pytest.raises(SystemExit):
        p.parse_args(["--output_dir", "out"])                      # --prompt_json is required
    f = tmp_path / "p.json"
    f.write_text(json.dumps({"a/b": "a cat", "c": {"text_prompt": "a dog", "image_prompt": "x.png"}, "d": {"prompt": "a fox"}}))
    tasks = g.load_tasks(str(f), None)
    assert [t["text_prompt"] for t in tasks] == ["a cat", "a dog", "a fox"]
    assert g.load_tasks(str(f), 2) == tasks[:2]
    gid, path = g.video_path_for(tmp_path, tasks[0], 0, 42)
    assert gid == "a_b" and path == tmp_path / "a_b" / "seed_42.mp4"
    f.write_text(json.dumps([{"group_id": 7, "text_prompt": "x"}, {"prompt": "y"}]))
    tasks = g.load_tasks(str(f), None)
    assert g.video_path_for(tmp_path, tasks[0], 0, 1)[0] == "7" and g.video_path_for(tmp_path, tasks[1], 1, 1)[0] == "1"
    f.write_text("3")
    assert g.load_tasks(str(f), None) is None


def test_generate_cli_i2v_and_1_5_surfaces(tmp_path):
    """Flag surfaces of the I2V and 1.5 generate CLIs (generate/CogVideoX-5B-I2V.py:100-112, generate/CogVideoX1.5-5B.py:102-111)
    and the I2V task / image resolution rules (host logic only)."""
    import json
    from videogpa_b200.generate import cogvideox1_5_5b as g15
    from videogpa_b200.generate import cogvideox_5b_i2v as gi
    a = gi.build_parser().parse_args(["--prompt_json", "p", "--output_dir", "o", "--base_dir", "imgs"])
    assert (a.base_model, a.base_dir, a.seed, a.num_inference_steps, a.guidance_scale, a.fps) == ("THUDM/CogVideoX-5B-I2V", "imgs", 42, 50, 6.0, 8)
    b = g15.build_parser().parse_args(["--prompt_json", "p", "--output_dir", "o"])
    assert (b.lora_weight, b.height, b.width, b.num_frames, b.fps) == (0.2, 768, 1360, 81, 16)
    f = tmp_path / "p.json"
    f.write_text(json.dumps({"k1": {"text_prompt": "a", "image_prompt": "a.png"}, "k2": {"prompt": "b", "image_path": "b.png"}}))
    tasks = gi.load_tasks(str(f), None)
    assert [t[0] for t in tasks] == ["k1", "k2"] and gi.load_tasks(str(f), 1) == tasks[:1]
    (tmp_path / "imgs").mkdir()
    (tmp_path / "imgs" / "a.png").write_bytes(b"x")
    assert gi.resolve_image(tasks[0][1], str(tmp_path / "imgs")) == str(tmp_path / "imgs" / "a.png")
    assert gi.resolve_image(tasks[1][1], None) == "b.png" and gi.resolve_image({}, None) == ""
    f.write_text(json.dumps([{"group_id": 3, "text_prompt": "x", "input_image_path": "c.png"}]))
    assert gi.load_tasks(str(f), None)[0][0] == 3
