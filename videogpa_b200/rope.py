"""3-D rotary position table of CogVideoX (diffusers get_3d_rotary_pos_embed).

Built once per pipeline call on the host in fp32 (17 550 x 64 values) and consumed by the fused QKV
GEMM epilogue. head_dim 64 splits into t/h/w = 16/24/24; theta = 10000; positions are the integer
latent grid at native resolution; cos/sin are repeat-interleaved over the rotation pairs.
"""
from __future__ import annotations

import torch


def get_3d_rotary_pos_embed(head_dim: int, grid_h: int, grid_w: int, num_frames: int, theta: float = 10000.0,
                            device="cpu"):
    dim_t, dim_h, dim_w = head_dim // 4, head_dim // 8 * 3, head_dim // 8 * 3

    def axis(n, dim):
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
        ang = torch.outer(torch.arange(n, dtype=torch.float32), freqs)
        return ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)

    (ct, st), (ch, sh), (cw, sw) = axis(num_frames, dim_t), axis(grid_h, dim_h), axis(grid_w, dim_w)

    def bcast(t, h, w):
        t = t[:, None, None, :].expand(-1, grid_h, grid_w, -1)
        h = h[None, :, None, :].expand(num_frames, -1, grid_w, -1)
        w = w[None, None, :, :].expand(num_frames, grid_h, -1, -1)
        return torch.cat([t, h, w], dim=-1).reshape(num_frames * grid_h * grid_w, head_dim).contiguous().to(device)

    return bcast(ct, ch, cw), bcast(st, sh, sw)
