"""CPU: host-side logic of the VAE decoder and Wan mirrors against the oracles / torch (no GPU, no kernels)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import vae_torch as V
from oracle import wan_torch as OW


def test_vae_frame_batches_and_time_maps_match_torch_interpolate():
    from videogpa_b200.vae import AutoencoderKLCogVideoXDecoder as D
    for n in (1, 2, 3, 5, 13, 21):
        assert D.frame_batches(n, 2) == V.frame_batches(n, 2)
    assert D.frame_batches(13, 2) == [(0, 3), (3, 5), (5, 7), (7, 9), (9, 11), (11, 13)]
    # SpatialNorm3D resizes zq with nearest interpolation, the first frame separately when T > 1 is odd (App. A.5)
    for T, Tz in [(1, 1), (2, 2), (3, 3), (4, 2), (5, 3), (8, 2), (9, 3), (2, 1), (4, 1), (16, 2)]:
        idx = torch.arange(Tz, dtype=torch.float32).view(1, 1, Tz, 1, 1)
        if T > 1 and T % 2 == 1:
            ref = torch.cat([F.interpolate(idx[:, :, :1], size=(1, 1, 1)), F.interpolate(idx[:, :, 1:], size=(T - 1, 1, 1))], dim=2) \
                if Tz > 1 else torch.zeros(1, 1, T, 1, 1)
        else:
            ref = F.interpolate(idx, size=(T, 1, 1))
        assert D._tz_map(T, Tz) == [int(v) for v in ref.flatten().tolist()], (T, Tz)


def test_vae_tiling_geometry_matches_oracle_and_survey():
    from videogpa_b200.vae import VAEDecoderConfig
    geo = V.tiling_geometry(V.VAEConfig())
    assert geo == dict(tile_latent_h=30, tile_latent_w=45, overlap_h=25, overlap_w=36, blend_h=40, blend_w=72, limit_h=200, limit_w=288)
    c = VAEDecoderConfig()
    assert (c.sample_height // 2 // 8, c.sample_width // 2 // 8) == (30, 45)
    assert int(30 * (1 - c.tile_overlap_factor_height)) == 25 and int(45 * (1 - c.tile_overlap_factor_width)) == 36
    assert list(range(0, 60, 25)) == [0, 25, 50] and list(range(0, 90, 36)) == [0, 36, 72]          # the 3 x 3 tiles of 60 x 90 latents


def test_wan_rope_sigmas_and_timestep_rules_match_oracle():
    from videogpa_b200.wan import WanConfig, WanTransformer3D, flow_sigmas, rope_tables
    for (dim, heads, f, h, w) in [(3072, 24, 3, 4, 5), (256, 2, 2, 3, 3)]:
        cos, sin = rope_tables(WanConfig(dim=dim, num_heads=heads), f, h, w)
        oc, os_ = OW.rope_tables(OW.WanConfig(dim=dim, num_heads=heads), f, h, w)
        assert torch.equal(cos, oc) and torch.equal(sin, os_)
        assert cos.shape == (f * h * w, 128)
    d = 128
    assert [d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)] == [44, 42, 42]                             # App. A.7 split
    s, so = flow_sigmas(50, 5.0), OW.flow_sigmas(50, 5.0)
    assert len(s) == 51 and max(abs(a - float(b)) for a, b in zip(s, so)) < 1e-12
    two = WanTransformer3D._two_timesteps
    assert two(torch.tensor(500.0), 10, 4) == (500.0, 500.0)
    t = torch.full((10,), 900.0); t[:4] = 0
    assert two(t, 10, 4) == (0.0, 900.0)
    with pytest.raises(RuntimeError):
        two(torch.arange(10.0), 10, 4)
    with pytest.raises(RuntimeError):
        two(torch.zeros(7), 10, 4)


def test_c_abi_argument_validation_of_new_entry_points(lib):
    """Validation happens before any CUDA call, so it is checkable without a GPU."""
    import ctypes as C
    from videogpa_b200 import _lib
    assert lib.vgpa_conv3d_causal_bf16(None, None) != 0 and b"null args" in lib.vgpa_last_error()
    a = _lib.Conv3dArgs()
    a.x = a.w = a.out = 0x1000
    a.T, a.H, a.W, a.Cin, a.Cout, a.Cout_pad, a.KT, a.ldo = 1, 8, 8, 48, 64, 64, 3, 64
    assert lib.vgpa_conv3d_causal_bf16(C.byref(a), None) != 0 and b"multiple of 64" in lib.vgpa_last_error()
    a.Cin, a.KT = 64, 2
    assert lib.vgpa_conv3d_causal_bf16(C.byref(a), None) != 0 and b"KT must be 1 or 3" in lib.vgpa_last_error()
    assert lib.vgpa_groupnorm_workspace_bytes(512) == 592 * 2 * 512 * 4
    assert lib.vgpa_groupnorm_stats_bf16(None, 10, 64, 32, 1e-6, None, 0, None, None) != 0
    assert lib.vgpa_spatialnorm_apply_bf16(None, None) != 0 and lib.vgpa_vae_compose_tiles_bf16(None, None) != 0
    assert lib.vgpa_rmsnorm_rope_bf16(None, 1, 256, 256, None, 1e-6, None, None, 128, 0, None) != 0
    assert lib.vgpa_add_rows_bf16(None, None, None, 1, 8, 8, None) != 0
    at = _lib.AttentionArgs()
    at.q = at.k = at.v = at.out = 0x1000
    at.B, at.H, at.Sq, at.Skv, at.head_dim = 1, 1, 8, 8, 96
    assert lib.vgpa_attention_bf16(C.byref(at), None) != 0 and b"head_dim must be 64 or 128" in lib.vgpa_last_error()


def test_encode_frame_selection_matches_reference_rule():
    """train/CogVideoX-5B/02_encode.py:55-63: short clips keep every frame, long clips use truncated linspace; frames are
    scaled to [0, 1] and laid out [3, F, H, W]."""
    import numpy as np
    from videogpa_b200.encode import frames_to_video_tensor, select_frame_indices
    assert select_frame_indices(10, 49).tolist() == list(range(10))
    assert select_frame_indices(49, 49).tolist() == list(range(49))
    idx = select_frame_indices(120, 49)
    assert idx.tolist() == np.linspace(0, 119, 49).astype(int).tolist() and idx[0] == 0 and idx[-1] == 119
    fr = np.random.default_rng(0).integers(0, 256, (60, 4, 6, 3), dtype=np.uint8)
    v = frames_to_video_tensor(fr, 49, device="cpu")
    assert v.shape == (3, 49, 4, 6) and float(v.max()) <= 1.0 and float(v.min()) >= 0.0
    assert torch.equal(v[:, 5], torch.from_numpy(fr[select_frame_indices(60, 49)[5]]).float().div(255).permute(2, 0, 1))
    with pytest.raises(RuntimeError):
        frames_to_video_tensor(np.zeros((4, 4, 4), dtype=np.uint8), 49, device="cpu")


def test_encoder_oracle_shapes_and_tiling_geometry():
    """Oracle-side invariants of the VAE encoder restatement: 49 frames -> passes of 9 + 5 x 8 -> 13 latent frames; full-size
    tiling geometry (sample tiles 240 x 360, stride 200 x 288, latent blend 5 / 9, crop 25 x 36)."""
    from oracle import vae_torch as V
    assert V.frame_batches(49, V.NUM_SAMPLE_FRAMES_BATCH_SIZE) == [(0, 9), (9, 17), (17, 25), (25, 33), (33, 41), (41, 49)]
    geo = V.encode_tiling_geometry(V.VAEConfig())
    assert geo == dict(tile_h=240, tile_w=360, overlap_h=200, overlap_w=288, blend_h=5, blend_w=9, limit_h=25, limit_w=36)
    cfg = V.VAEConfig(block_out_channels=(32, 32, 32, 32), layers_per_block=1, sample_height=32, sample_width=32)
    sd = V.random_encoder_state_dict(cfg, seed=1)
    x = torch.rand(1, 3, 9, 16, 16, generator=torch.Generator().manual_seed(0))
    m = V.encode(sd, cfg, x, tiling=False)
    assert m.shape == (1, 32, 3, 2, 2)
    # causal across passes: a 17-frame clip is encoded as passes of 9 + 8 frames, so its first 3 latent frames equal the
    # 9-frame clip's (GroupNorm statistics are per pass, conv caches only flow forward)
    x17 = torch.cat([x, torch.rand(1, 3, 8, 16, 16, generator=torch.Generator().manual_seed(1))], dim=2)
    m17 = V.encode(sd, cfg, x17, tiling=False)
    assert m17.shape == (1, 32, 5, 2, 2) and torch.allclose(m17[:, :, :3], m, atol=1e-6)


def test_t5_position_buckets_match_transformers():
    """The relative-position bucket rule is integer work: bit-exact against transformers' own static method (the library the
    reference's text_encoder comes from), for the encoder's bidirectional case at 226 and 512 tokens."""
    from transformers.models.t5.modeling_t5 import T5Attention
    from videogpa_b200.t5 import T5Config, position_buckets
    for S in (1, 17, 226, 512):
        ctx = torch.arange(S)[:, None]
        mem = torch.arange(S)[None, :]
        ref = T5Attention._relative_position_bucket(mem - ctx, bidirectional=True, num_buckets=32, max_distance=128)
        got = position_buckets(S, 32, 128)
        assert got.dtype == torch.long and torch.equal(got, ref)
        assert int(got.min()) >= 0 and int(got.max()) < 32
    c = T5Config()
    assert (c.d_model, c.d_kv, c.num_heads, c.d_ff, c.num_layers) == (4096, 64, 64, 10240, 24)


def test_t5_abi_argument_validation():
    from videogpa_b200 import _lib
    L = _lib.load()
    assert L.vgpa_t5_attention_bf16(None, None, None, None, None, 1, 1, 8, 64, 64, None) != 0
    assert b"null" in L.vgpa_last_error()
    assert L.vgpa_t5_attention_bf16(16, 16, 16, 16, 16, 1, 1, 513, 64, 64, None) != 0 and b"512" in L.vgpa_last_error()
    assert L.vgpa_t5_attention_bf16(16, 16, 16, 16, 16, 1, 2, 8, 64, 128, None) != 0
    assert L.vgpa_gated_mul_bf16(16, 16, 16, 4, 12, 16, 16, 16, None) != 0 and b"multiple of 8" in L.vgpa_last_error()
    assert L.vgpa_gated_mul_bf16(16, 16, 16, 4, 16, 8, 16, 16, None) != 0


def test_wan_generate_cli_surface(tmp_path):
    """generate/Wan2.2-TI2V-5B.py:140-153 flags and defaults, task parsing (:77-89) and the latent grid of the Wan2.2 VAE
    strides (81 frames 704x1280 -> 21 x 44 x 80 latents -> 18 480 tokens, BASELINE.json configs[3])."""
    import json
    from videogpa_b200.generate import wan2_2_ti2v_5b as g
    p = g.build_parser()
    a = p.parse_args(["--model_path", "m", "--prompt_json", "p.json", "--output_dir", "o"])
    assert (a.lora_path, a.lora_weight, a.base_dir, a.gpu_id, a.seed, a.num_prompts, a.frame_num, a.shift, a.sampling_steps,
            a.guide_scale, a.fps) == (None, 0.2, None, 0, 42, None, 81, 5.0, 50, 5.0, 24)
    with pytest.raises(SystemExit):
        p.parse_args(["--prompt_json", "p.json", "--output_dir", "o"])          # --model_path is required
    assert g.latent_grid(81, 704, 1280) == (21, 44, 80, 18480)
    f = tmp_path / "p.json"
    f.write_text(json.dumps({"a/b": {"text_prompt": "x", "image_prompt": "i.png"}, "c": {"prompt": "y", "image_path": "j.png"}}))
    t = g.load_tasks(str(f), None)
    assert [k for k, _ in t] == ["a/b", "c"] and g.load_tasks(str(f), 1) == t[:1]
    f.write_text(json.dumps([{"group_id": 7, "text_prompt": "x"}, {"text_prompt": "y"}]))
    assert [k for k, _ in g.load_tasks(str(f), None)] == [7, 1]
    f.write_text(json.dumps("nope"))
    assert g.load_tasks(str(f), None) is None


def test_flow_unipc_scheduler_orders_and_schedule():
    """schedulers.FlowUniPCMultistepScheduler (WanTI2V.generate's default sampler) on a linear flow with the exact solution
    x(s) = x(s0) (s / s0)^(1 - a): UniPC-p is of order p + 1, so halving the step divides the error by ~4 (p = 1) and ~8
    (p = 2); the last step (sigma -> 0) returns the data prediction; the shifted schedule matches its closed form."""
    import math
    import numpy as np
    from videogpa_b200.schedulers import FlowUniPCMultistepScheduler
    a = 0.3

    def err(N, order):
        sch = FlowUniPCMultistepScheduler(solver_order=order, lower_order_final=False)
        sch.set_timesteps(N, sigmas=np.exp(np.linspace(math.log(0.95), math.log(0.05), N + 1)))
        x = torch.ones(2)
        for i in range(N):
            x = sch.step(x * (1 - a) / sch.sigmas[i], x)
        return abs(float(x[0]) - (0.05 / 0.95) ** (1 - a))

    e1 = [err(N, 1) for N in (8, 16, 32)]
    e2 = [err(N, 2) for N in (8, 16, 32)]
    assert 3.5 < e1[0] / e1[1] < 4.5 and 3.5 < e1[1] / e1[2] < 4.5
    assert 6.0 < e2[0] / e2[1] < 10.0 and 6.0 < e2[1] / e2[2] < 10.0
    assert e2[2] < e1[2] / 30
    sch = FlowUniPCMultistepScheduler()
    sch.set_timesteps(50, shift=5.0)
    lin = np.linspace(0.999, 0.0, 51)[:-1]
    assert np.allclose(sch.sigmas[:-1], 5.0 * lin / (1 + 4.0 * lin)) and sch.sigmas[-1] == 0.0 and len(sch.timesteps) == 50
    assert abs(sch.timesteps[0] - 1000 * 5.0 * 0.999 / (1 + 4.0 * 0.999)) < 1e-9
    x = torch.full((3,), 2.0)
    for i in range(50):
        x = sch.step(x * (1 - a) / sch.sigmas[i], x)          # perfect-model limit: the final sample is the last x0 prediction
    assert torch.isfinite(x).all() and float(x[0]) > 0
    with pytest.raises(RuntimeError):
        FlowUniPCMultistepScheduler().step(torch.zeros(1), torch.zeros(1))
    with pytest.raises(RuntimeError):
        sch.set_timesteps(3, sigmas=[0.5, 0.6, 0.2, 0.1])


def test_wan_umt5_checkpoint_name_mapping():
    """t5.wan_umt5_to_transformers_names: Wan's own umT5 parameter names -> transformers' UMT5EncoderModel names (the mapping is
    recalled, so the test pins its behaviour, not the Wan repository): every rule, and a loud failure on an unknown key."""
    from videogpa_b200.t5 import wan_umt5_to_transformers_names as f
    src = {"token_embedding.weight": 0, "norm.weight": 1, "blocks.0.norm1.weight": 2, "blocks.0.norm2.weight": 3, "blocks.23.attn.q.weight": 4,
           "blocks.23.attn.k.weight": 5, "blocks.23.attn.v.weight": 6, "blocks.23.attn.o.weight": 7, "blocks.7.pos_embedding.embedding.weight": 8,
           "blocks.7.ffn.gate.0.weight": 9, "blocks.7.ffn.fc1.weight": 10, "blocks.7.ffn.fc2.weight": 11}
    out = f(src)
    assert out == {"shared.weight": 0, "encoder.final_layer_norm.weight": 1, "encoder.block.0.layer.0.layer_norm.weight": 2,
                   "encoder.block.0.layer.1.layer_norm.weight": 3, "encoder.block.23.layer.0.SelfAttention.q.weight": 4,
                   "encoder.block.23.layer.0.SelfAttention.k.weight": 5, "encoder.block.23.layer.0.SelfAttention.v.weight": 6,
                   "encoder.block.23.layer.0.SelfAttention.o.weight": 7,
                   "encoder.block.7.layer.0.SelfAttention.relative_attention_bias.weight": 8,
                   "encoder.block.7.layer.1.DenseReluDense.wi_0.weight": 9, "encoder.block.7.layer.1.DenseReluDense.wi_1.weight": 10,
                   "encoder.block.7.layer.1.DenseReluDense.wo.weight": 11}
    with pytest.raises(RuntimeError, match="unexpected parameter"):
        f({"blocks.0.attn.q.bias": 0})
    from pathlib import Path
    from videogpa_b200.generate.wan2_2_ti2v_5b import WanPromptEncoder
    assert WanPromptEncoder.available(Path("/nonexistent")) is False
