"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the Wan2.2 TI2V-5B DiT forward — SURVEY.md §8 row a-16.

Restates `WanModel.forward` of the external Wan2.2 repository as the reference calls it
(`generate/Wan2.2-TI2V-5B.py:120-129` via `WanTI2V.generate`; `train/Wan2.2-TI2V-5B/03_train.py:228-233`; dims in
`train/Wan2.2-TI2V-5B/03_train.py:9-13`). Wan2.2 is NOT vendored in /root/reference and no commit is pinned anywhere, so
this follows SURVEY.md App. A.7 from recollection of `wan/modules/model.py`: **parity unpinned**. Only tests/,
__graft_entry__.smoke() and bench.py's CPU arm may import it.

Model: dim 3072, ffn 14336, freq_dim 256, 24 heads x 128, 30 layers, in = out = 48, text_len 512, patch (1,2,2), RMSNorm
q/k over the full dim, cross-attention LayerNorm with affine (norm3), eps 1e-6. Block:
    e = modulation[1,6,D] + time_projection(SiLU(time_embedding(sinusoid_256(t))))        per token in 2.2
    x += self_attn(LN(x) (1 + e1) + e0) e2 ;  x += cross_attn(LN_affine(x), context) ;  x += ffn(LN(x) (1 + e4) + e3) e5
RoPE: complex pairs, head_dim 128 split 44/42/42 over (t, h, w), theta 10000.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class WanConfig:
    dim: int = 3072
    ffn_dim: int = 14336
    freq_dim: int = 256
    num_heads: int = 24
    num_layers: int = 30
    in_dim: int = 48
    out_dim: int = 48
    text_dim: int = 4096
    text_len: int = 512
    patch_size: tuple = (1, 2, 2)
    eps: float = 1e-6

    @property
    def head_dim(self):
        return self.dim // self.num_heads


def random_state_dict(cfg: WanConfig, seed: int = 21, std: float = 0.02, dtype=torch.float32, device=None) -> dict:
    """device: generate on that device (its own generator stream), for the full-size GPU checks."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    class _R:                                   # torch.randn on `device` with the shared generator
        @staticmethod
        def randn(*shape, generator=None):
            return torch.randn(*shape, generator=g, device=device)

    def lin(name, o, i, s=std):
        sd[name + ".weight"] = (_R.randn(o, i) * s).to(dtype)
        sd[name + ".bias"] = (_R.randn(o) * 0.02).to(dtype)

    D = cfg.dim
    sd["patch_embedding.weight"] = (_R.randn(D, cfg.in_dim, *cfg.patch_size) * std).to(dtype)
    sd["patch_embedding.bias"] = (_R.randn(D) * 0.02).to(dtype)
    lin("text_embedding.0", D, cfg.text_dim); lin("text_embedding.2", D, D)
    lin("time_embedding.0", D, cfg.freq_dim); lin("time_embedding.2", D, D)
    lin("time_projection.1", 6 * D, D)
    for i in range(cfg.num_layers):
        b = f"blocks.{i}."
        for a in ("self_attn", "cross_attn"):
            for m in ("q", "k", "v", "o"):
                lin(b + f"{a}.{m}", D, D)
            sd[b + f"{a}.norm_q.weight"] = (1 + 0.1 * _R.randn(D)).to(dtype)
            sd[b + f"{a}.norm_k.weight"] = (1 + 0.1 * _R.randn(D)).to(dtype)
        sd[b + "norm3.weight"] = (1 + 0.1 * _R.randn(D)).to(dtype)
        sd[b + "norm3.bias"] = (0.05 * _R.randn(D)).to(dtype)
        lin(b + "ffn.0", cfg.ffn_dim, D); lin(b + "ffn.2", D, cfg.ffn_dim)
        sd[b + "modulation"] = (_R.randn(1, 6, D) / D ** 0.5).to(dtype)
    lin("head.head", cfg.out_dim * math.prod(cfg.patch_size), D)
    sd["head.modulation"] = (_R.randn(1, 2, D) / D ** 0.5).to(dtype)
    return sd


def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    half = dim // 2
    position = position.to(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half, dtype=torch.float64, device=position.device) / half))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def rope_tables(cfg: WanConfig, F_: int, H: int, W: int, device=None):
    """(cos, sin) [F*H*W, head_dim] fp32, repeat-interleaved over the complex pairs (rope_params + rope_apply)."""
    d = cfg.head_dim
    dims = [d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)]

    def ang(n, dim):
        return torch.outer(torch.arange(n, dtype=torch.float64, device=device),
                           1.0 / torch.pow(10000, torch.arange(0, dim, 2, dtype=torch.float64, device=device) / dim))

    at, ah, aw = ang(F_, dims[0]), ang(H, dims[1]), ang(W, dims[2])
    a = torch.cat([at[:, None, None, :].expand(F_, H, W, -1), ah[None, :, None, :].expand(F_, H, W, -1),
                   aw[None, None, :, :].expand(F_, H, W, -1)], dim=-1).reshape(F_ * H * W, d // 2)
    return a.cos().repeat_interleave(2, dim=1).float(), a.sin().repeat_interleave(2, dim=1).float()


def rope_apply(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [S, heads, head_dim]: complex multiply on interleaved pairs."""
    xr, xi = x.double()[..., 0::2], x.double()[..., 1::2]
    c, s = cos.double()[:, None, 0::2], sin.double()[:, None, 0::2]
    out = torch.stack([xr * c - xi * s, xi * c + xr * s], dim=-1).flatten(-2)
    return out.float()


def rms_norm(x, w, eps):
    return (x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype) * w


def layer_norm(x, eps, w=None, b=None):
    return F.layer_norm(x.float(), x.shape[-1:], w, b, eps).to(x.dtype)


def attention(q, k, v):
    """q [Sq, n, d], k/v [Skv, n, d] -> [Sq, n*d]"""
    from .dit_torch import sdpa                      # explicit fp32 softmax for fp32 CUDA inputs, F.sdpa otherwise
    o = sdpa(q.transpose(0, 1)[None], k.transpose(0, 1)[None], v.transpose(0, 1)[None])[0]
    return o.transpose(0, 1).flatten(1)


def block_forward(sd, cfg: WanConfig, i: int, x, e0, context, cos, sin):
    """x [S, D]; e0 [S, 6, D] (time projection per token); context [text_len, D] (embedded)."""
    p = f"blocks.{i}."
    n, d = cfg.num_heads, cfg.head_dim
    lin = lambda name, t: F.linear(t, sd[p + name + ".weight"], sd[p + name + ".bias"])
    e = (sd[p + "modulation"][0][None] + e0).unbind(1)                      # 6 x [S, D]
    h = layer_norm(x, cfg.eps) * (1 + e[1]) + e[0]
    q = rms_norm(lin("self_attn.q", h), sd[p + "self_attn.norm_q.weight"], cfg.eps).view(-1, n, d)
    k = rms_norm(lin("self_attn.k", h), sd[p + "self_attn.norm_k.weight"], cfg.eps).view(-1, n, d)
    v = lin("self_attn.v", h).view(-1, n, d)
    y = lin("self_attn.o", attention(rope_apply(q, cos, sin), rope_apply(k, cos, sin), v))
    x = x + y * e[2]
    h = layer_norm(x, cfg.eps, sd[p + "norm3.weight"], sd[p + "norm3.bias"])
    q = rms_norm(lin("cross_attn.q", h), sd[p + "cross_attn.norm_q.weight"], cfg.eps).view(-1, n, d)
    k = rms_norm(lin("cross_attn.k", context), sd[p + "cross_attn.norm_k.weight"], cfg.eps).view(-1, n, d)
    v = lin("cross_attn.v", context).view(-1, n, d)
    x = x + lin("cross_attn.o", attention(q, k, v))
    h = layer_norm(x, cfg.eps) * (1 + e[4]) + e[3]
    y = lin("ffn.2", F.gelu(lin("ffn.0", h), approximate="tanh"))
    return x + y * e[5]


def model_forward(sd, cfg: WanConfig, x, t, context, num_layers: int | None = None):
    """x [C, F, H, W] latent; t [S] per-token timesteps (or a scalar); context [L <= text_len, text_dim] -> [C, F, H, W]."""
    C, F_, H, W = x.shape
    pt, ph, pw = cfg.patch_size
    f, h, w = F_ // pt, H // ph, W // pw
    S = f * h * w
    tok = F.conv3d(x[None], sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=cfg.patch_size)[0]   # [D, f, h, w]
    xs = tok.flatten(1).transpose(0, 1)                                                                           # [S, D]
    t = torch.as_tensor(t, dtype=torch.float32, device=x.device).reshape(-1)
    if t.numel() == 1:
        t = t.expand(S)
    emb = sinusoidal_embedding_1d(cfg.freq_dim, t).float()
    e = F.linear(F.silu(F.linear(emb, sd["time_embedding.0.weight"], sd["time_embedding.0.bias"])),
                 sd["time_embedding.2.weight"], sd["time_embedding.2.bias"])                                      # [S, D]
    e0 = F.linear(F.silu(e), sd["time_projection.1.weight"], sd["time_projection.1.bias"]).unflatten(1, (6, cfg.dim))
    ctx = torch.cat([context, context.new_zeros(cfg.text_len - context.shape[0], context.shape[1])])
    ctx = F.linear(F.gelu(F.linear(ctx, sd["text_embedding.0.weight"], sd["text_embedding.0.bias"]), approximate="tanh"),
                   sd["text_embedding.2.weight"], sd["text_embedding.2.bias"])
    cos, sin = rope_tables(cfg, f, h, w, device=x.device)
    L = cfg.num_layers if num_layers is None else num_layers
    for i in range(L):
        xs = block_forward(sd, cfg, i, xs, e0, ctx, cos, sin)
    em = (sd["head.modulation"][0][None] + e[:, None]).unbind(1)
    y = F.linear(layer_norm(xs, cfg.eps) * (1 + em[1]) + em[0], sd["head.head.weight"], sd["head.head.bias"])      # [S, pt*ph*pw*C]
    u = y.view(f, h, w, pt, ph, pw, cfg.out_dim)
    u = torch.einsum("fhwpqrc->cfphqwr", u)
    return u.reshape(cfg.out_dim, f * pt, h * ph, w * pw)


def flow_sigmas(num_steps: int, shift: float = 5.0, num_train_timesteps: int = 1000):
    """Shifted flow-matching sigma schedule of the Wan samplers: sigma' = shift * s / (1 + (shift - 1) * s)."""
    s = torch.linspace(1.0, 1.0 / num_train_timesteps, num_steps, dtype=torch.float64)
    s = shift * s / (1 + (shift - 1) * s)
    return torch.cat([s, s.new_zeros(1)])


def euler_flow_step(x, v, sigma, sigma_next):
    return x + (sigma_next - sigma) * v
