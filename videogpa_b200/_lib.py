"""ctypes binding of libvideogpa_b200.so (the C-ABI declared in include/videogpa_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is
raised (the reference CLIs catch ``Exception`` per item and continue, generate/CogVideoX-5B.py:79).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libvideogpa_b200.so"
_lib = None

c_void_p = C.c_void_p
c_int = C.c_int
c_i32 = C.c_int32
c_i64 = C.c_int64
c_float = C.c_float
c_double = C.c_double
c_fp = C.POINTER(C.c_float)


class LinearArgs(C.Structure):
    _fields_ = [
        ("A", c_void_p), ("W", c_void_p), ("bias", c_void_p), ("out", c_void_p),
        ("M", c_i32), ("N", c_i32), ("K", c_i32), ("lda", c_i32), ("ldo", c_i32),
        ("epilogue", c_i32), ("rows_per_sample", c_i32), ("text_rows", c_i32),
        ("gate_txt", c_void_p), ("gate_vid", c_void_p), ("gate_stride_b", c_i64),
        ("ln_q_w", c_void_p), ("ln_q_b", c_void_p), ("ln_k_w", c_void_p), ("ln_k_b", c_void_p),
        ("ln_eps", c_float),
        ("rope_cos", c_void_p), ("rope_sin", c_void_p),
        ("model_dim", c_i32),
        ("alpha", c_float),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("out", c_void_p),
        ("B", c_i32), ("H", c_i32), ("Sq", c_i32), ("Skv", c_i32), ("head_dim", c_i32),
        ("scale", c_float),
        ("q_row_stride", c_i64), ("k_row_stride", c_i64), ("v_row_stride", c_i64), ("out_row_stride", c_i64),
        ("q_batch_stride", c_i64), ("k_batch_stride", c_i64), ("v_batch_stride", c_i64), ("out_batch_stride", c_i64),
        ("lse", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class AttentionBwdArgs(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("out", c_void_p), ("d_out", c_void_p), ("lse", c_void_p),
        ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
        ("B", c_i32), ("H", c_i32), ("Sq", c_i32), ("Skv", c_i32), ("head_dim", c_i32),
        ("scale", c_float),
        ("q_row_stride", c_i64), ("k_row_stride", c_i64), ("v_row_stride", c_i64), ("out_row_stride", c_i64),
        ("dout_row_stride", c_i64), ("dq_row_stride", c_i64), ("dk_row_stride", c_i64), ("dv_row_stride", c_i64),
        ("q_batch_stride", c_i64), ("k_batch_stride", c_i64), ("v_batch_stride", c_i64), ("out_batch_stride", c_i64),
        ("dout_batch_stride", c_i64), ("dq_batch_stride", c_i64), ("dk_batch_stride", c_i64), ("dv_batch_stride", c_i64),
        ("workspace", c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class LayerNormArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("out", c_void_p),
        ("rows", c_i32), ("D", c_i32),
        ("ldx", c_i64), ("ldo", c_i64),
        ("ln_weight", c_void_p), ("ln_bias", c_void_p),
        ("eps", c_float),
        ("rows_per_sample", c_i32), ("text_rows", c_i32),
        ("shift_txt", c_void_p), ("scale_txt", c_void_p), ("shift_vid", c_void_p), ("scale_vid", c_void_p),
        ("mod_stride_b", c_i64),
        ("x_is_f32", c_i32),
    ]


class DpoArgs(C.Structure):
    _fields_ = [
        ("tensors", c_void_p * 6), ("is_bf16", c_i32 * 6),
        ("B", c_i32), ("n_per_sample", c_i64),
        ("beta", c_float), ("label_smoothing", c_float), ("loss_type", c_i32),
        ("d_out5", c_void_p), ("d_err4", c_void_p), ("d_coef", c_void_p),
        ("d_workspace", c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class SchedArgs(C.Structure):
    _fields_ = [
        ("pred_uncond", c_void_p), ("pred_cond", c_void_p), ("sample", c_void_p), ("prev_sample", c_void_p),
        ("x0_old", c_void_p), ("x0_out", c_void_p), ("noise", c_void_p),
        ("n", c_i64), ("mode", c_i32),
        ("guidance", c_float), ("sqrt_alpha_t", c_float), ("sqrt_beta_t", c_float),
        ("c_sample", c_float), ("c_x0", c_float), ("c_x0_old", c_float), ("c_noise", c_float),
    ]


class Conv3dArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("w", c_void_p), ("bias", c_void_p), ("residual", c_void_p), ("out", c_void_p),
        ("T", c_i32), ("H", c_i32), ("W", c_i32), ("Cin", c_i32), ("Cout", c_i32), ("Cout_pad", c_i32), ("KT", c_i32),
        ("ldo", c_i32), ("ld_res", c_i32),
        ("gn_mean_rstd", c_void_p), ("gn_workspace", c_void_p), ("gn_groups", c_i32), ("gn_eps", C.c_float),
    ]


class SpatialNormArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("out", c_void_p), ("mean_rstd", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
        ("y_lat", c_void_p), ("b_lat", c_void_p), ("ld_lat", c_i64),
        ("T", c_i32), ("H", c_i32), ("W", c_i32), ("C", c_i32), ("groups", c_i32),
        ("Hz", c_i32), ("Wz", c_i32), ("shift", c_i32),
        ("tz_of_t", c_i32 * 16),
        ("silu", c_i32),
    ]


class ComposeArgs(C.Structure):
    _fields_ = [
        ("tiles", c_void_p * 16),
        ("rows", c_i32), ("cols", c_i32),
        ("th", c_i32 * 4), ("tw", c_i32 * 4),
        ("T", c_i32), ("H", c_i32), ("W", c_i32), ("ldc", c_i32),
        ("blend_h", c_i32), ("blend_w", c_i32), ("limit_h", c_i32), ("limit_w", c_i32),
        ("out", c_void_p),
    ]


EPI_BIAS, EPI_BIAS_GELU, EPI_GATE_RES, EPI_QKV, EPI_ACCUM, EPI_GATE_RES_F32 = 0, 1, 2, 3, 4, 5
ACT_NONE, ACT_SILU = 0, 1
SCHED_DDIM, SCHED_DPM = 0, 1

# name -> (restype, argtypes); every symbol include/videogpa_b200.h declares must be listed here
# (tests/test_abi.py cross-checks this table against the header).
SIGNATURES = {
    "vgpa_last_error": (C.c_char_p, []),
    "vgpa_abi_version": (c_int, []),
    "vgpa_device_sm_count": (c_int, []),
    "vgpa_linear_bf16": (c_int, [C.POINTER(LinearArgs), c_void_p]),
    "vgpa_attention_bf16": (c_int, [C.POINTER(AttentionArgs), c_void_p]),
    "vgpa_attention_workspace_bytes": (C.c_size_t, [c_int, c_int, c_int]),
    "vgpa_attention_bwd_workspace_bytes": (C.c_size_t, [c_int, c_int, c_int]),
    "vgpa_attention_bwd_bf16": (c_int, [C.POINTER(AttentionBwdArgs), c_void_p]),
    "vgpa_layernorm_modulate_bwd_bf16": (c_int, [C.POINTER(LayerNormArgs), c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p]),
    "vgpa_head_layernorm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_float, c_int, c_void_p]),
    "vgpa_gelu_tanh_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "vgpa_transpose_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_i64, c_i64, c_void_p]),
    "vgpa_scale_cols_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_i64, c_i64, c_i64, c_int, c_int, c_void_p, c_void_p, c_i64, c_void_p]),
    "vgpa_layernorm_modulate_bf16": (c_int, [C.POINTER(LayerNormArgs), c_void_p]),
    "vgpa_linear_smallm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_i64, c_i64, c_int, c_void_p]),
    "vgpa_timestep_embedding_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "vgpa_patchify_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vgpa_unpatchify_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_i64, c_void_p]),
    "vgpa_cfg_scheduler_step": (c_int, [C.POINTER(SchedArgs), c_void_p]),
    "vgpa_mvcs_workspace_bytes": (C.c_size_t, [c_int, c_int, c_int, c_int]),
    "vgpa_mvcs_blocks_per_pair": (c_int, [c_int, c_int, c_int, c_int]),
    "vgpa_mvcs_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                C.c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vgpa_reproject_workspace_bytes": (C.c_size_t, [c_int, c_int, c_int]),
    "vgpa_reproject_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_void_p,
                                     C.c_size_t, c_void_p, c_void_p]),
    "vgpa_pointcloud_workspace_bytes": (C.c_size_t, [c_i64]),
    "vgpa_pointcloud_filter": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_double, c_void_p, C.c_size_t,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vgpa_epipolar_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vgpa_motion_score": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "vgpa_mse_workspace_bytes": (C.c_size_t, []),
    "vgpa_mse_range_normalized": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_i64, c_int, c_int,
                                          c_int, c_void_p, C.c_size_t, c_void_p, c_void_p]),
    "vgpa_unproject_depth": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vgpa_dpo_workspace_bytes": (C.c_size_t, [c_int, c_i64]),
    "vgpa_dpo_loss_forward": (c_int, [C.POINTER(DpoArgs), c_void_p]),
    "vgpa_dpo_loss_backward": (c_int, [C.POINTER(DpoArgs), c_void_p, c_void_p, c_void_p, c_void_p]),
    "vgpa_rmsnorm_rope_bf16": (c_int, [c_void_p, c_int, c_int, c_i64, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "vgpa_add_rows_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_i64, c_i64, c_void_p]),
    "vgpa_t5_attention_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_i64, c_i64, c_void_p]),
    "vgpa_gated_mul_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_i64, c_i64, c_i64, c_void_p]),
    "vgpa_conv3d_causal_bf16": (c_int, [C.POINTER(Conv3dArgs), c_void_p]),
    "vgpa_conv3d_gn_workspace_bytes": (C.c_size_t, [c_int]),
    "vgpa_groupnorm_workspace_bytes": (C.c_size_t, [c_int]),
    "vgpa_groupnorm_stats_bf16": (c_int, [c_void_p, c_i64, c_int, c_int, c_float, c_void_p, C.c_size_t, c_void_p, c_void_p]),
    "vgpa_spatialnorm_apply_bf16": (c_int, [C.POINTER(SpatialNormArgs), c_void_p]),
    "vgpa_upsample_nearest_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, C.POINTER(c_i32), c_void_p]),
    "vgpa_vae_compose_tiles_bf16": (c_int, [C.POINTER(ComposeArgs), c_void_p]),
}


def lib_path() -> Path:
    return _LIB_PATH


ABI_VERSION = 4          # VGPA_ABI_VERSION of include/videogpa_b200.h


def load():
    """Load the library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m videogpa_b200.build` "
            "(there is no CPU or PyTorch fallback for the CUDA hot path)")
    lib = C.CDLL(os.fspath(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.vgpa_abi_version() != ABI_VERSION:           # a stale in-tree .so would silently mis-read the argument structs
        raise RuntimeError(f"{_LIB_PATH} has ABI version {lib.vgpa_abi_version()}, the bindings expect {ABI_VERSION}: "
                           "rebuild it with `python -m videogpa_b200.build`")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vgpa_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode(errors='replace') if msg else ''}")


def ptr(t) -> int | None:
    """Device/host address of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
