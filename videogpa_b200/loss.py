"""DPO loss on the fused sm_100a kernel — drop-in for the reference's train/loss.py (same
DPOLoss(beta, label_smoothing, loss_type) constructor, forward(6 tensors) -> LossOutput,
create_loss_strategy(strategy, beta, label_smoothing)); differentiable w.r.t. v_win and v_lose
through a fused backward kernel (03_train.py:157 calls it inside the Lightning training step).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import _lib
from ._lib import DpoArgs


@dataclass
class LossOutput:
    """train/loss.py:15-22."""
    loss: torch.Tensor
    reward_margin: torch.Tensor
    winner_reward: torch.Tensor
    loser_reward: torch.Tensor
    accuracy: torch.Tensor


def _prep(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.float()
    return t.contiguous()


def _make_args(tensors, beta, label_smoothing, loss_type, out5, err4, coef, ws) -> DpoArgs:
    a = DpoArgs()
    for i, t in enumerate(tensors):
        a.tensors[i] = t.data_ptr()
        a.is_bf16[i] = 1 if t.dtype == torch.bfloat16 else 0
    a.B = tensors[0].shape[0]
    a.n_per_sample = tensors[0].numel() // tensors[0].shape[0]
    a.beta, a.label_smoothing, a.loss_type = beta, label_smoothing, loss_type
    a.d_out5, a.d_err4, a.d_coef = out5.data_ptr(), _lib.ptr(err4), coef.data_ptr()
    a.d_workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    return a


class _DPOFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v_win, v_lose, v_win_ref, v_lose_ref, v_win_target, v_lose_target, beta, label_smoothing, loss_type):
        lib = _lib.load()
        ts = [_prep(t, n) for t, n in zip((v_win, v_lose, v_win_ref, v_lose_ref, v_win_target, v_lose_target),
                                          ("v_win", "v_lose", "v_win_ref", "v_lose_ref", "v_win_target", "v_lose_target"))]
        shape = ts[0].shape
        for t in ts[1:]:
            if t.shape != shape:
                raise RuntimeError("all six DPO tensors must have the same shape")
        B = shape[0]
        n = ts[0].numel() // B
        dev = ts[0].device
        out5 = torch.empty(5, dtype=torch.float32, device=dev)
        err4 = torch.empty((4, B), dtype=torch.float32, device=dev)
        coef = torch.empty((2, B), dtype=torch.float32, device=dev)
        ws = torch.empty(((lib.vgpa_dpo_workspace_bytes(B, n) + 255) // 256) * 256, dtype=torch.uint8, device=dev)
        args = _make_args(ts, beta, label_smoothing, loss_type, out5, err4, coef, ws)
        _lib.check(lib.vgpa_dpo_loss_forward(C.byref(args), _lib.current_stream()), "vgpa_dpo_loss_forward")
        ctx.save_for_backward(ts[0], ts[1], ts[4], ts[5], coef)
        ctx.meta = (beta, label_smoothing, loss_type, v_win.dtype, v_lose.dtype)
        ctx.mark_non_differentiable(err4)
        return out5[0], out5[1], out5[2], out5[3], out5[4], err4

    @staticmethod
    def backward(ctx, g_loss, g_margin, g_wr, g_lr, g_acc, g_err):
        lib = _lib.load()
        v_win, v_lose, t_win, t_lose, coef = ctx.saved_tensors
        beta, ls, lt, dt_w, dt_l = ctx.meta
        # only the loss is a training signal in the reference (03_train.py:161-169 logs the rest detached)
        g = g_loss.to(torch.float32).contiguous().reshape(1)
        gw, gl = torch.empty_like(v_win), torch.empty_like(v_lose)
        dummy = torch.empty(5, dtype=torch.float32, device=v_win.device)
        ws = torch.empty(256, dtype=torch.uint8, device=v_win.device)
        args = _make_args([v_win, v_lose, v_win, v_lose, t_win, t_lose], beta, ls, lt, dummy, None, coef, ws)
        _lib.check(lib.vgpa_dpo_loss_backward(C.byref(args), g.data_ptr(), gw.data_ptr(), gl.data_ptr(), _lib.current_stream()),
                   "vgpa_dpo_loss_backward")
        return gw.to(dt_w), gl.to(dt_l), None, None, None, None, None, None, None


class DPOLoss(nn.Module):
    """train/loss.py:25-121."""

    def __init__(self, beta: float = 500.0, label_smoothing: float = 0.0, loss_type: str = "sigmoid"):
        super().__init__()
        self.beta = beta
        self.label_smoothing = label_smoothing
        self.loss_type = loss_type

    def forward(self, v_win, v_lose, v_win_ref, v_lose_ref, v_win_target, v_lose_target) -> LossOutput:
        if self.loss_type not in ("sigmoid", "hinge"):
            raise ValueError(f"Unknown loss type: {self.loss_type}")
        lt = 0 if self.loss_type == "sigmoid" else 1
        loss, margin, wr, lr, acc, _ = _DPOFunction.apply(v_win, v_lose, v_win_ref, v_lose_ref, v_win_target, v_lose_target,
                                                          float(self.beta), float(self.label_smoothing), lt)
        return LossOutput(loss=loss, reward_margin=margin.detach(), winner_reward=wr.detach(), loser_reward=lr.detach(),
                          accuracy=acc.detach())


class SFTLoss(nn.Module):
    """`strategy="sft"` of create_loss_strategy (train/loss.py:139-152): F.mse_loss(v_pred, v_target) through
    the same kernel (loss type VGPA_DPO_SFT; the prediction fills the winner slot)."""

    def forward(self, v_pred, v_target, **kwargs):
        loss, _, _, _, _, _ = _DPOFunction.apply(v_pred, v_pred, v_pred, v_pred, v_target, v_target, 1.0, 0.0, 2)
        z = torch.tensor(0.0)
        return LossOutput(loss=loss, reward_margin=z, winner_reward=z, loser_reward=z, accuracy=z)


def create_loss_strategy(strategy: str = "dpo", beta: float = 1.0, label_smoothing: float = 0.0) -> nn.Module:
    """train/loss.py:124-155."""
    if strategy == "dpo":
        return DPOLoss(beta=beta, label_smoothing=label_smoothing)
    if strategy == "sft":
        return SFTLoss()
    raise ValueError(f"Unknown strategy: {strategy}")
