"""Host-side wrappers of the dense sm_100a kernels (GEMM / attention / norm / embed) over the C-ABI.

These are the building blocks the DiT mirror (videogpa_b200/transformer.py) strings together; each
wrapper only validates shapes/dtypes, allocates the output on the current CUDA stream and calls the
C entry point. No torch math happens here and there is no fallback: a missing library or a CPU
tensor raises RuntimeError.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (AttentionBwdArgs, ACT_NONE, ACT_SILU, EPI_ACCUM, EPI_BIAS, EPI_BIAS_GELU, EPI_GATE_RES, EPI_GATE_RES_F32, EPI_QKV,
                   SCHED_DDIM, SCHED_DPM, AttentionArgs, LayerNormArgs, LinearArgs, SchedArgs)

BF16 = torch.bfloat16


def _req(t: torch.Tensor, dtype, name: str, contiguous: bool = True) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t


def _rows(t: torch.Tensor, name: str) -> torch.Tensor:
    """2-D view whose last dim is contiguous (row stride may exceed the width)."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"{name} must be 2-D with a contiguous last dimension")
    return t


def linear(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *,
           out: torch.Tensor | None = None, epilogue: int = EPI_BIAS,
           rows_per_sample: int = 0, text_rows: int = 0,
           gate_txt: torch.Tensor | None = None, gate_vid: torch.Tensor | None = None,
           gate_stride_b: int = 0,
           ln_q: tuple[torch.Tensor, torch.Tensor] | None = None,
           ln_k: tuple[torch.Tensor, torch.Tensor] | None = None, ln_eps: float = 1e-6,
           rope: tuple[torch.Tensor, torch.Tensor] | None = None, model_dim: int = 0,
           alpha: float = 1.0) -> torch.Tensor:
    """out = epilogue(a[M,K] @ w[N,K]^T); see vgpa_linear_bf16 in include/videogpa_b200.h."""
    lib = _lib.load()
    _rows(_req(a, BF16, "a", contiguous=False), "a")
    _req(w, BF16, "w")
    M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise RuntimeError(f"linear: K mismatch {K} vs {K2}")
    if out is None:
        if epilogue in (EPI_GATE_RES, EPI_ACCUM, EPI_GATE_RES_F32):
            raise RuntimeError("linear: this epilogue updates `out` in place; pass the tensor to update")
        out = torch.empty((M, N), dtype=BF16, device=a.device)
    _rows(_req(out, torch.float32 if epilogue == EPI_GATE_RES_F32 else BF16, "out", contiguous=False), "out")
    if out.shape[0] != M or out.shape[1] != N:
        raise RuntimeError(f"linear: out shape {tuple(out.shape)} != ({M}, {N})")
    args = LinearArgs()
    args.A, args.W, args.bias, args.out = a.data_ptr(), w.data_ptr(), _lib.ptr(bias), out.data_ptr()
    args.M, args.N, args.K, args.lda, args.ldo = M, N, K, a.stride(0), out.stride(0)
    args.epilogue = epilogue
    args.rows_per_sample, args.text_rows = rows_per_sample, text_rows
    args.gate_txt, args.gate_vid, args.gate_stride_b = _lib.ptr(gate_txt), _lib.ptr(gate_vid), gate_stride_b
    if ln_q is not None:
        args.ln_q_w = _req(ln_q[0], torch.float32, "ln_q.w").data_ptr()
        args.ln_q_b = _req(ln_q[1], torch.float32, "ln_q.b").data_ptr()
    if ln_k is not None:
        args.ln_k_w = _req(ln_k[0], torch.float32, "ln_k.w").data_ptr()
        args.ln_k_b = _req(ln_k[1], torch.float32, "ln_k.b").data_ptr()
    args.ln_eps = ln_eps
    if rope is not None:
        args.rope_cos = _req(rope[0], torch.float32, "rope.cos").data_ptr()
        args.rope_sin = _req(rope[1], torch.float32, "rope.sin").data_ptr()
    args.model_dim = model_dim
    args.alpha = alpha
    _lib.check(lib.vgpa_linear_bf16(C.byref(args), _lib.current_stream()), "vgpa_linear_bf16")
    return out


_ATTN_WS: dict = {}


def _attention_workspace(lib, B: int, heads: int, head_dim: int, device) -> torch.Tensor | None:
    """Per-(device, stream) scratch for the |q|, |k| bounds of the bounded-softmax forward (head_dim 64)."""
    n = lib.vgpa_attention_workspace_bytes(B, heads, head_dim)
    if n == 0:
        return None
    if torch.cuda.is_current_stream_capturing():       # a captured launch keeps the pointer: give the graph its own buffer
        return torch.empty(max(n, 4096), dtype=torch.uint8, device=device)
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _ATTN_WS.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.empty(max(n, 4096), dtype=torch.uint8, device=device)
        _ATTN_WS[key] = ws
    return ws


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, *,
              out: torch.Tensor | None = None, scale: float = 0.0, head_dim: int = 64,
              lse: torch.Tensor | None = None, exact: bool = False) -> torch.Tensor:
    """softmax(q k^T * scale) v with heads packed along the last dim.

    q: [B, Sq, >=heads*64] view, k/v: [B, Skv, >=heads*64] views (last dim contiguous; they may be
    column slices of one fused qkv buffer). Returns [B, Sq, heads*64] bf16.
    exact=True withholds the workspace, which forces the online-softmax kernel for every head.
    """
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, BF16, n, contiguous=False)
        if t.dim() != 3 or t.stride(2) != 1:
            raise RuntimeError(f"attention: {n} must be [B, S, cols] with a contiguous last dimension")
    B, Sq, _ = q.shape
    Skv = k.shape[1]
    if out is None:
        out = torch.empty((B, Sq, heads * head_dim), dtype=BF16, device=q.device)
    _req(out, BF16, "out", contiguous=False)
    a = AttentionArgs()
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.B, a.H, a.Sq, a.Skv, a.head_dim = B, heads, Sq, Skv, head_dim
    a.scale = scale
    a.q_row_stride, a.k_row_stride, a.v_row_stride, a.out_row_stride = q.stride(1), k.stride(1), v.stride(1), out.stride(1)
    a.q_batch_stride, a.k_batch_stride, a.v_batch_stride, a.out_batch_stride = q.stride(0), k.stride(0), v.stride(0), out.stride(0)
    if lse is not None:                                   # [B, heads, Sq] fp32, saved for attention_backward
        _req(lse, torch.float32, "lse")
        if tuple(lse.shape) != (B, heads, Sq):
            raise RuntimeError(f"attention: lse must be [{B}, {heads}, {Sq}]")
        a.lse = lse.data_ptr()
    ws = None if exact else _attention_workspace(lib, B, heads, head_dim, q.device)
    if ws is not None:
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    _lib.check(lib.vgpa_attention_bf16(C.byref(a), _lib.current_stream()), "vgpa_attention_bf16")
    return out


def attention_backward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, d_out: torch.Tensor,
                       lse: torch.Tensor, heads: int, *, scale: float = 0.0, grads=None):
    """(dq, dk, dv) of `attention` (head_dim 64) from the saved output and logsumexp; see vgpa_attention_bwd_bf16."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out"), (d_out, "d_out")):
        _req(t, BF16, n, contiguous=False)
        if t.dim() != 3 or t.stride(2) != 1:
            raise RuntimeError(f"attention_backward: {n} must be [B, S, cols] with a contiguous last dimension")
    _req(lse, torch.float32, "lse")
    B, Sq, _ = q.shape
    Skv = k.shape[1]
    if tuple(lse.shape) != (B, heads, Sq):
        raise RuntimeError(f"attention_backward: lse must be [{B}, {heads}, {Sq}]")
    if grads is not None:                                 # caller-provided (dq, dk, dv), e.g. column slices of one buffer
        dq, dk, dv = grads
        for t, n in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
            _req(t, BF16, n, contiguous=False)
            if t.dim() != 3 or t.stride(2) != 1:
                raise RuntimeError(f"attention_backward: {n} must be [B, S, cols] with a contiguous last dimension")
    else:
        dq = torch.empty((B, Sq, heads * 64), dtype=BF16, device=q.device)
        dk = torch.empty((B, Skv, heads * 64), dtype=BF16, device=q.device)
        dv = torch.empty((B, Skv, heads * 64), dtype=BF16, device=q.device)
    ws = torch.empty(lib.vgpa_attention_bwd_workspace_bytes(B, heads, Sq), dtype=torch.uint8, device=q.device)
    a = AttentionBwdArgs()
    a.q, a.k, a.v, a.out, a.d_out, a.lse = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), d_out.data_ptr(), lse.data_ptr()
    a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.B, a.H, a.Sq, a.Skv, a.head_dim = B, heads, Sq, Skv, 64
    a.scale = scale
    a.q_row_stride, a.k_row_stride, a.v_row_stride, a.out_row_stride = q.stride(1), k.stride(1), v.stride(1), out.stride(1)
    a.dout_row_stride, a.dq_row_stride, a.dk_row_stride, a.dv_row_stride = d_out.stride(1), dq.stride(1), dk.stride(1), dv.stride(1)
    a.q_batch_stride, a.k_batch_stride, a.v_batch_stride, a.out_batch_stride = q.stride(0), k.stride(0), v.stride(0), out.stride(0)
    a.dout_batch_stride, a.dq_batch_stride, a.dk_batch_stride, a.dv_batch_stride = d_out.stride(0), dq.stride(0), dk.stride(0), dv.stride(0)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    _lib.check(lib.vgpa_attention_bwd_bf16(C.byref(a), _lib.current_stream()), "vgpa_attention_bwd_bf16")
    return dq, dk, dv


def layernorm_modulate(x: torch.Tensor, ln_weight: torch.Tensor | None, ln_bias: torch.Tensor | None, *,
                       eps: float = 1e-5, out: torch.Tensor | None = None,
                       rows_per_sample: int = 0, text_rows: int = 0,
                       shift_txt=None, scale_txt=None, shift_vid=None, scale_vid=None,
                       mod_stride_b: int = 0) -> torch.Tensor:
    """LN(x) * (1 + scale[b, seg]) + shift[b, seg] over [rows, D]; see vgpa_layernorm_modulate_bf16."""
    lib = _lib.load()
    x_f32 = x.dtype == torch.float32                     # fp32 residual stream (Wan2.2): fp32 math, one bf16 rounding
    _rows(_req(x, torch.float32 if x_f32 else BF16, "x", contiguous=False), "x")
    rows, D = x.shape
    if out is None:
        out = torch.empty((rows, D), dtype=BF16, device=x.device)
    _rows(_req(out, BF16, "out", contiguous=False), "out")
    a = LayerNormArgs()
    a.x, a.out, a.rows, a.D, a.ldx, a.ldo = x.data_ptr(), out.data_ptr(), rows, D, x.stride(0), out.stride(0)
    a.ln_weight, a.ln_bias, a.eps = _lib.ptr(ln_weight), _lib.ptr(ln_bias), eps
    a.rows_per_sample, a.text_rows = rows_per_sample, text_rows
    a.shift_txt, a.scale_txt, a.shift_vid, a.scale_vid = (_lib.ptr(shift_txt), _lib.ptr(scale_txt),
                                                        _lib.ptr(shift_vid), _lib.ptr(scale_vid))
    a.mod_stride_b = mod_stride_b
    a.x_is_f32 = 1 if x_f32 else 0
    _lib.check(lib.vgpa_layernorm_modulate_bf16(C.byref(a), _lib.current_stream()), "vgpa_layernorm_modulate_bf16")
    return out


def linear_smallm(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, *, act_in: int = ACT_NONE) -> torch.Tensor:
    """out[M<=8, N] = bias + act_in(x) @ w^T (conditioning path GEMV)."""
    lib = _lib.load()
    _rows(_req(x, BF16, "x", contiguous=False), "x")
    _req(w, BF16, "w")
    M, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise RuntimeError(f"linear_smallm: K mismatch {K} vs {w.shape[1]}")
    out = torch.empty((M, N), dtype=BF16, device=x.device)
    _lib.check(lib.vgpa_linear_smallm_bf16(x.data_ptr(), w.data_ptr(), _lib.ptr(bias), out.data_ptr(), M, N, K,
                                           x.stride(0), out.stride(0), act_in, _lib.current_stream()),
               "vgpa_linear_smallm_bf16")
    return out


def timestep_embedding(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    lib = _lib.load()
    _req(timesteps, torch.float32, "timesteps")
    B = timesteps.numel()
    out = torch.empty((B, dim), dtype=BF16, device=timesteps.device)
    _lib.check(lib.vgpa_timestep_embedding_bf16(timesteps.data_ptr(), out.data_ptr(), B, dim, _lib.current_stream()),
               "vgpa_timestep_embedding_bf16")
    return out


def patchify(x: torch.Tensor) -> torch.Tensor:
    """[BF, C, H, W] -> [BF*(H/2)*(W/2), C*4]."""
    lib = _lib.load()
    _req(x, BF16, "x")
    BF_, Cc, H, W = x.shape
    out = torch.empty((BF_ * (H // 2) * (W // 2), Cc * 4), dtype=BF16, device=x.device)
    _lib.check(lib.vgpa_patchify_bf16(x.data_ptr(), out.data_ptr(), BF_, Cc, H, W, _lib.current_stream()), "vgpa_patchify_bf16")
    return out


def unpatchify(tok: torch.Tensor, BF_: int, Cc: int, H: int, W: int) -> torch.Tensor:
    """[BF*(H/2)*(W/2), >=C*4] -> [BF, C, H, W]."""
    lib = _lib.load()
    _rows(_req(tok, BF16, "tok", contiguous=False), "tok")
    out = torch.empty((BF_, Cc, H, W), dtype=BF16, device=tok.device)
    _lib.check(lib.vgpa_unpatchify_bf16(tok.data_ptr(), out.data_ptr(), BF_, Cc, H, W, tok.stride(0), _lib.current_stream()),
               "vgpa_unpatchify_bf16")
    return out


def cfg_scheduler_step(pred_cond: torch.Tensor, pred_uncond: torch.Tensor | None, sample: torch.Tensor, *,
                       mode: int, guidance: float, sqrt_alpha_t: float, sqrt_beta_t: float,
                       c_sample: float, c_x0: float, c_x0_old: float = 0.0, c_noise: float = 0.0,
                       x0_old: torch.Tensor | None = None, x0_out: torch.Tensor | None = None,
                       noise: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    lib = _lib.load()
    _req(pred_cond, BF16, "pred_cond")
    _req(sample, BF16, "sample")
    if pred_uncond is not None:
        _req(pred_uncond, BF16, "pred_uncond")
    if out is None:
        out = torch.empty_like(sample)
    a = SchedArgs()
    a.pred_uncond, a.pred_cond, a.sample, a.prev_sample = _lib.ptr(pred_uncond), pred_cond.data_ptr(), sample.data_ptr(), out.data_ptr()
    a.x0_old = _lib.ptr(_req(x0_old, torch.float32, "x0_old")) if x0_old is not None else None
    a.x0_out = _lib.ptr(_req(x0_out, torch.float32, "x0_out")) if x0_out is not None else None
    a.noise = _lib.ptr(_req(noise, BF16, "noise")) if noise is not None else None
    a.n, a.mode = sample.numel(), mode
    a.guidance, a.sqrt_alpha_t, a.sqrt_beta_t = guidance, sqrt_alpha_t, sqrt_beta_t
    a.c_sample, a.c_x0, a.c_x0_old, a.c_noise = c_sample, c_x0, c_x0_old, c_noise
    _lib.check(lib.vgpa_cfg_scheduler_step(C.byref(a), _lib.current_stream()), "vgpa_cfg_scheduler_step")
    return out
