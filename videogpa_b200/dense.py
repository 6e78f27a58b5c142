"""Host-side wrappers of the dense sm_100a kernels (GEMM / attention / norm) over the C-ABI.

These are the building blocks the DiT mirror (videogpa_b200/transformer.py) strings together; each
wrapper only validates shapes/dtypes, allocates the output on the current CUDA stream and calls the
C entry point. No torch math happens here.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import EPI_BIAS, EPI_BIAS_GELU, EPI_GATE_RES, EPI_QKV, LinearArgs


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t


def linear(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *,
           out: torch.Tensor | None = None, epilogue: int = EPI_BIAS,
           rows_per_sample: int = 0, text_rows: int = 0,
           gate_txt: torch.Tensor | None = None, gate_vid: torch.Tensor | None = None,
           gate_stride_b: int = 0,
           ln_q: tuple[torch.Tensor, torch.Tensor] | None = None,
           ln_k: tuple[torch.Tensor, torch.Tensor] | None = None, ln_eps: float = 1e-6,
           rope: tuple[torch.Tensor, torch.Tensor] | None = None, model_dim: int = 0) -> torch.Tensor:
    """out = epilogue(a[M,K] @ w[N,K]^T); see vgpa_linear_bf16 in include/videogpa_b200.h."""
    lib = _lib.load()
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise RuntimeError(f"linear: K mismatch {K} vs {K2}")
    if out is None:
        if epilogue == EPI_GATE_RES:
            raise RuntimeError("linear: EPI_GATE_RES updates `out` in place; pass the residual stream")
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _req(out, torch.bfloat16, "out")
    args = LinearArgs()
    args.A, args.W, args.bias, args.out = a.data_ptr(), w.data_ptr(), _lib.ptr(bias), out.data_ptr()
    args.M, args.N, args.K, args.lda, args.ldo = M, N, K, a.stride(0), out.stride(0)
    args.epilogue = epilogue
    args.rows_per_sample, args.text_rows = rows_per_sample, text_rows
    args.gate_txt, args.gate_vid, args.gate_stride_b = _lib.ptr(gate_txt), _lib.ptr(gate_vid), gate_stride_b
    if ln_q is not None:
        args.ln_q_w, args.ln_q_b = _req(ln_q[0], torch.float32, "ln_q.w").data_ptr(), _req(ln_q[1], torch.float32, "ln_q.b").data_ptr()
    if ln_k is not None:
        args.ln_k_w, args.ln_k_b = _req(ln_k[0], torch.float32, "ln_k.w").data_ptr(), _req(ln_k[1], torch.float32, "ln_k.b").data_ptr()
    args.ln_eps = ln_eps
    if rope is not None:
        args.rope_cos = _req(rope[0], torch.float32, "rope.cos").data_ptr()
        args.rope_sin = _req(rope[1], torch.float32, "rope.sin").data_ptr()
    args.model_dim = model_dim
    _lib.check(lib.vgpa_linear_bf16(C.byref(args), _lib.current_stream()), "vgpa_linear_bf16")
    return out
