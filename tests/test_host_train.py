"""CPU checks for the training-step row: C-ABI argument validation of the backward entry points (no
compute without a GPU), PEFT parameter naming, and the product path refusing CPU tensors."""
import ctypes as C

import pytest
import torch

from videogpa_b200 import _lib


def test_backward_entry_points_validate_arguments():
    L = _lib.load()
    a = _lib.AttentionBwdArgs()
    assert L.vgpa_attention_bwd_bf16(None, None) != 0 and b"null" in L.vgpa_last_error()
    a.head_dim = 128
    assert L.vgpa_attention_bwd_bf16(C.byref(a), None) != 0 and b"head_dim" in L.vgpa_last_error()
    a.head_dim, a.B, a.H, a.Sq, a.Skv = 64, 1, 2, 16, 16
    assert L.vgpa_attention_bwd_bf16(C.byref(a), None) != 0 and b"null tensor" in L.vgpa_last_error()
    for f in ("q", "k", "v", "out", "d_out", "lse", "dq", "dk", "dv"):
        setattr(a, f, 256)
    assert L.vgpa_attention_bwd_bf16(C.byref(a), None) != 0 and b"workspace" in L.vgpa_last_error()
    assert L.vgpa_attention_bwd_workspace_bytes(2, 48, 17776) == 2 * 48 * 17776 * 4
    a.workspace, a.workspace_bytes = 256, 1 << 20
    assert L.vgpa_attention_bwd_bf16(C.byref(a), None) != 0 and b"row strides" in L.vgpa_last_error()     # strides are 0
    # forward: the logsumexp output exists for head_dim 64 only
    f = _lib.AttentionArgs()
    f.q = f.k = f.v = f.out = 256
    f.lse = 256
    f.B, f.H, f.Sq, f.Skv, f.head_dim = 1, 2, 16, 16, 128
    for n in ("q_row_stride", "k_row_stride", "v_row_stride", "out_row_stride"):
        setattr(f, n, 256)
    assert L.vgpa_attention_bf16(C.byref(f), None) != 0 and b"logsumexp" in L.vgpa_last_error()

    ln = _lib.LayerNormArgs()
    assert L.vgpa_layernorm_modulate_bwd_bf16(C.byref(ln), None, 0, None, 0, None, 0, None) != 0
    ln.x, ln.rows, ln.D, ln.ldx = 256, 4, 100, 100
    assert L.vgpa_layernorm_modulate_bwd_bf16(C.byref(ln), 256, 100, None, 0, 256, 100, None) != 0 and b"multiple of 256" in L.vgpa_last_error()
    assert L.vgpa_head_layernorm_bf16(256, None, 256, 4, 2, 128, 0, 128, 256, None, 1e-6, 0, None) != 0 and b"bias is null" in L.vgpa_last_error()
    assert L.vgpa_head_layernorm_bf16(256, None, 256, 4, 2, 128, 0, 128, 256, 256, 1e-6, 1, None) != 0 and b"dy is null" in L.vgpa_last_error()
    assert L.vgpa_head_layernorm_bf16(256, 256, 256, 4, 2, 64, 128, 128, 256, 256, 1e-6, 1, None) != 0        # ldx < heads * 64
    assert L.vgpa_gelu_tanh_bf16(256, None, 256, 12, 0, None) != 0 and b"multiple of 8" in L.vgpa_last_error()
    assert L.vgpa_gelu_tanh_bf16(256, None, 256, 16, 1, None) != 0                                            # backward without dy
    assert L.vgpa_scale_cols_bf16(256, None, 256, 4, 64, 64, 0, 64, 0, 0, None, None, 0, None) != 0 and b"null" in L.vgpa_last_error()
    assert L.vgpa_scale_cols_bf16(256, 256, 256, 4, 64, 64, 8, 64, 0, 0, 256, 256, 64, None) != 0 and b"ld_add" in L.vgpa_last_error()


def test_training_path_refuses_cpu_tensors_and_bad_rank():
    from videogpa_b200 import dense
    from videogpa_b200.train_dit import LoRATrainableTransformer, TARGETS
    assert TARGETS == ("to_q", "to_k", "to_v", "to_out.0")         # checkpoints/*/adapter_config.json target_modules
    q = torch.zeros(1, 16, 128, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        dense.attention_backward(q, q, q, q, q, torch.zeros(1, 2, 16), 2)

    class _Base:                                                    # the rank check happens before any device work
        config = None
        device = torch.device("cpu")
        blocks = []
    with pytest.raises(RuntimeError):
        LoRATrainableTransformer(_Base(), r=16)


def test_dpo_step_requires_trainable_for_training():
    from videogpa_b200.train_step import DPOSharedStep

    class _T:
        device = torch.device("cpu")
    step = DPOSharedStep(_T(), _T())
    with pytest.raises(RuntimeError):
        step.training_step({})


def test_train_cli_config_and_schedule(tmp_path):
    """train/CogVideoX-5B/03_train.py: DEFAULT_CONFIG values (:39-81), YAML `training:` override and --devices parsing
    (:290-305), the cosine-with-warm-up multiplier of diffusers' get_cosine_schedule_with_warmup, the 98/2 split (:237-242)."""
    import math
    from videogpa_b200.train import cogvideox_5b as t
    c = t.DEFAULT_CONFIG
    assert (c["learning_rate"], c["beta"], c["max_steps"], c["warmup_steps"], c["batch_size"], c["accumulate_grad_batches"],
            c["gradient_clip_val"], c["lora_rank"], c["lora_alpha"], c["min_gap"], c["metric_mode"]) == \
        (5e-6, 1.0, 10000, 500, 1, 2, 1.0, 64, 128.0, 0.05, "min")
    y = tmp_path / "cfg.yaml"
    y.write_text("training:\n  max_steps: 7\n  learning_rate: 1.0e-4\nother:\n  x: 1\n")
    cfg = t.load_config(t.build_parser().parse_args(["--config", str(y), "--devices", "2,3", "--base_path", "/data"]))
    assert cfg["max_steps"] == 7 and cfg["learning_rate"] == 1e-4 and cfg["devices"] == [2, 3] and cfg["base_path"] == "/data"
    assert cfg["warmup_steps"] == 500                                  # untouched keys keep the defaults
    f = t.cosine_schedule_with_warmup
    assert f(0, 500, 10000) == 0.0 and f(250, 500, 10000) == 0.5 and f(500, 500, 10000) == 1.0
    assert abs(f(5250, 500, 10000) - 0.5) < 1e-12 and abs(f(10000, 500, 10000)) < 1e-12
    assert abs(f(2875, 500, 10000) - 0.5 * (1 + math.cos(math.pi * 0.25))) < 1e-12
    tr, va = t.split_dataset(list(range(100)))
    tr2, _ = t.split_dataset(list(range(100)))
    assert len(tr) == 98 and len(va) == 2 and list(tr.indices) == list(tr2.indices)
