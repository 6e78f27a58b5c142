"""torch restatement of the CogVideoX denoiser math (test infrastructure; PARITY UNPINNED, see oracle/__init__.py).

Follows SURVEY.md App. A.1-A.4, which restates diffusers>=0.31:
  CogVideoXTransformer3DModel.forward, CogVideoXBlock, CogVideoXLayerNormZero, AdaLayerNorm,
  CogVideoXAttnProcessor2_0, get_3d_rotary_pos_embed / apply_rotary_emb, CogVideoXDDIMScheduler,
  CogVideoXDPMScheduler, and peft's LoRA merge. Reference call sites: generate/CogVideoX-5B.py:17-31,72-77,
  train/CogVideoX-5B/03_train.py:101-157.
The forward runs in whatever dtype the state dict is in: fp32 for the accuracy oracle, bf16 to
reproduce eager-bf16 rounding points. State-dict names are diffusers' (App. A.1).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class DiTConfig:
    num_attention_heads: int = 48
    attention_head_dim: int = 64
    in_channels: int = 16
    out_channels: int = 16
    time_embed_dim: int = 512
    text_embed_dim: int = 4096
    num_layers: int = 42
    patch_size: int = 2
    patch_size_t: int | None = None      # CogVideoX1.5: 2 (Linear patch embed over (c, pt, ph, pw), no patch bias)
    sample_width: int = 90
    sample_height: int = 60
    sample_frames: int = 49
    temporal_compression_ratio: int = 4
    max_text_seq_length: int = 226
    norm_eps: float = 1e-5
    use_rotary_positional_embeddings: bool = True
    use_learned_positional_embeddings: bool = False
    ffn_mult: int = 4

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


def random_state_dict(cfg: DiTConfig, seed: int = 1234, dtype=torch.float32, std: float = 0.02,
                      randomize_norms: bool = False, device=None) -> dict:
    """SURVEY.md §8d synthetic weights: N(0, 0.02^2) for Linear/Conv, LayerNorm gamma 1 beta 0, biases 0.

    randomize_norms=True also perturbs norm affine params and biases so parity tests exercise them.
    device: generate on that device (its own generator stream: CPU and CUDA draws differ), for the full-size GPU checks.
    """
    device = torch.device(device) if device is not None else torch.device("cpu")
    g = torch.Generator(device=device).manual_seed(seed)
    D, Tm = cfg.inner_dim, cfg.time_embed_dim
    sd = {}

    def randn(*shape):
        return torch.randn(*shape, generator=g, device=device)

    def lin(name, out_f, in_f, bias=True):
        sd[name + ".weight"] = randn(out_f, in_f) * std
        if bias:
            sd[name + ".bias"] = (randn(out_f) * std) if randomize_norms else torch.zeros(out_f, device=device)

    def norm(name, n):
        sd[name + ".weight"] = (1.0 + 0.1 * randn(n)) if randomize_norms else torch.ones(n, device=device)
        sd[name + ".bias"] = (0.05 * randn(n)) if randomize_norms else torch.zeros(n, device=device)

    p = cfg.patch_size
    if cfg.patch_size_t is None:
        sd["patch_embed.proj.weight"] = randn(D, cfg.in_channels, p, p) * std
    else:
        sd["patch_embed.proj.weight"] = randn(D, cfg.in_channels * cfg.patch_size_t * p * p) * std
    sd["patch_embed.proj.bias"] = (randn(D) * std) if randomize_norms else torch.zeros(D, device=device)
    lin("patch_embed.text_proj", D, cfg.text_embed_dim)
    if cfg.use_learned_positional_embeddings:
        n_tok = cfg.max_text_seq_length + ((cfg.sample_frames - 1) // cfg.temporal_compression_ratio + 1) * \
            (cfg.sample_height // p) * (cfg.sample_width // p)
        sd["patch_embed.pos_embedding"] = randn(1, n_tok, D) * std
    lin("time_embedding.linear_1", Tm, D)
    lin("time_embedding.linear_2", Tm, Tm)
    for i in range(cfg.num_layers):
        b = f"transformer_blocks.{i}."
        lin(b + "norm1.linear", 6 * D, Tm); norm(b + "norm1.norm", D)
        lin(b + "norm2.linear", 6 * D, Tm); norm(b + "norm2.norm", D)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(b + "attn1." + n, D, D)
        norm(b + "attn1.norm_q", cfg.attention_head_dim); norm(b + "attn1.norm_k", cfg.attention_head_dim)
        lin(b + "ff.net.0.proj", cfg.ffn_mult * D, D)
        lin(b + "ff.net.2", D, cfg.ffn_mult * D)
    norm("norm_final", D)
    lin("norm_out.linear", 2 * D, Tm); norm("norm_out.norm", D)
    lin("proj_out", p * p * (cfg.patch_size_t or 1) * cfg.out_channels, D)
    return {k: v.to(dtype) for k, v in sd.items()}


# ------------------------------------------------------------------ App. A.3: 3-D RoPE
def rope_3d(cfg: DiTConfig, num_frames: int, height: int, width: int):
    """(cos, sin) [F*h*w, head_dim] fp32 for latent grid (num_frames, height/p, width/p) at native resolution."""
    d = cfg.attention_head_dim
    dim_t, dim_h, dim_w = d // 4, d // 8 * 3, d // 8 * 3
    gh, gw = height // cfg.patch_size, width // cfg.patch_size

    def axis(n, dim):
        freqs = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
        ang = torch.outer(torch.arange(n, dtype=torch.float32), freqs)
        return ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)

    ct, st = axis(num_frames, dim_t)
    ch, sh = axis(gh, dim_h)
    cw, sw = axis(gw, dim_w)

    def bcast(t, h, w):
        t = t[:, None, None, :].expand(-1, gh, gw, -1)
        h = h[None, :, None, :].expand(num_frames, -1, gw, -1)
        w = w[None, None, :, :].expand(num_frames, gh, -1, -1)
        return torch.cat([t, h, w], dim=-1).reshape(num_frames * gh * gw, d).contiguous()

    return bcast(ct, ch, cw), bcast(st, sh, sw)


def apply_rotary(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [B, H, S, d]; interleaved-pair rotation in fp32, cast back (App. A.3)."""
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


def timestep_embedding(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)          # flip_sin_to_cos=True


# ------------------------------------------------------------------ App. A.1 / A.2
def sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """F.scaled_dot_product_attention. For fp32 CUDA inputs (the full-size checks run this oracle on the GPU) the softmax is
    written out over chunks of heads and queries, so that it is plain fp32 arithmetic with no fused-kernel precision choices
    and bounded memory (the caller switches TF32 off)."""
    if not (q.is_cuda and q.dtype == torch.float32):
        return F.scaled_dot_product_attention(q, k, v)
    B, Hh, S, d = q.shape
    out = torch.empty_like(q)
    rows = max(1, min(S, (1 << 28) // max(1, k.shape[2])))           # <= 1 GiB of fp32 scores per chunk
    for b in range(B):
        for h in range(Hh):
            kt = k[b, h].t()
            for r0 in range(0, S, rows):
                p = torch.softmax((q[b, h, r0:r0 + rows] @ kt) * (d ** -0.5), dim=-1)
                out[b, h, r0:r0 + rows] = p @ v[b, h]
    return out


def attention_block(sd, prefix, cfg, n_hs, n_enc, rope):
    Hh, d = cfg.num_attention_heads, cfg.attention_head_dim
    St = n_enc.shape[1]
    x = torch.cat([n_enc, n_hs], dim=1)
    B, S, _ = x.shape
    q = F.linear(x, sd[prefix + "to_q.weight"], sd[prefix + "to_q.bias"]).view(B, S, Hh, d).transpose(1, 2)
    k = F.linear(x, sd[prefix + "to_k.weight"], sd[prefix + "to_k.bias"]).view(B, S, Hh, d).transpose(1, 2)
    v = F.linear(x, sd[prefix + "to_v.weight"], sd[prefix + "to_v.bias"]).view(B, S, Hh, d).transpose(1, 2)
    q = F.layer_norm(q, (d,), sd[prefix + "norm_q.weight"], sd[prefix + "norm_q.bias"], 1e-6)
    k = F.layer_norm(k, (d,), sd[prefix + "norm_k.weight"], sd[prefix + "norm_k.bias"], 1e-6)
    if rope is not None:
        cos, sin = rope
        q = torch.cat([q[:, :, :St], apply_rotary(q[:, :, St:], cos, sin)], dim=2)
        k = torch.cat([k[:, :, :St], apply_rotary(k[:, :, St:], cos, sin)], dim=2)
    o = sdpa(q, k, v)
    o = o.transpose(1, 2).reshape(B, S, Hh * d)
    o = F.linear(o, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"])
    return o[:, St:], o[:, :St]


def layer_norm_zero(sd, prefix, cfg, hs, enc, emb):
    D = cfg.inner_dim
    mod = F.linear(F.silu(emb), sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])
    shift, scale, gate, enc_shift, enc_scale, enc_gate = mod.chunk(6, dim=1)
    w, b = sd[prefix + "norm.weight"], sd[prefix + "norm.bias"]
    n_hs = F.layer_norm(hs, (D,), w, b, cfg.norm_eps) * (1 + scale)[:, None, :] + shift[:, None, :]
    n_enc = F.layer_norm(enc, (D,), w, b, cfg.norm_eps) * (1 + enc_scale)[:, None, :] + enc_shift[:, None, :]
    return n_hs, n_enc, gate[:, None, :], enc_gate[:, None, :]


def transformer_forward(sd: dict, cfg: DiTConfig, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor,
                        timestep: torch.Tensor, image_rotary_emb=None, num_layers: int | None = None) -> torch.Tensor:
    """hidden_states [B, F, C, H, W], encoder_hidden_states [B, St, 4096], timestep [B] -> sample [B, F, C_out, H, W]."""
    dtype = sd["proj_out.weight"].dtype
    B, Fr, C, H, W = hidden_states.shape
    p, D = cfg.patch_size, cfg.inner_dim
    St = encoder_hidden_states.shape[1]
    t_emb = timestep_embedding(timestep, D).to(dtype)
    emb = F.linear(t_emb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
    emb = F.linear(F.silu(emb), sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])
    pt = cfg.patch_size_t
    if pt is None:
        x_img = F.conv2d(hidden_states.reshape(B * Fr, C, H, W).to(dtype), sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=p)
        x_img = x_img.view(B, Fr, D, H // p, W // p).flatten(3).transpose(2, 3).flatten(1, 2)
    else:                                                         # CogVideoXPatchEmbed with patch_size_t (1.5)
        e = hidden_states.to(dtype).reshape(B, Fr // pt, pt, C, H // p, p, W // p, p)
        e = e.permute(0, 1, 4, 6, 3, 2, 5, 7).flatten(4, 7).flatten(1, 3)
        x_img = F.linear(e, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"])
    x_txt = F.linear(encoder_hidden_states.to(dtype), sd["patch_embed.text_proj.weight"], sd["patch_embed.text_proj.bias"])
    x = torch.cat([x_txt, x_img], dim=1)
    if cfg.use_learned_positional_embeddings:
        x = x + sd["patch_embed.pos_embedding"][:, : x.shape[1]]
    enc, hs = x[:, :St], x[:, St:]
    rope = None
    if image_rotary_emb is not None:
        rope = (image_rotary_emb[0].float(), image_rotary_emb[1].float())
    L = cfg.num_layers if num_layers is None else num_layers
    for i in range(L):
        b = f"transformer_blocks.{i}."
        n_hs, n_enc, gate, enc_gate = layer_norm_zero(sd, b + "norm1.", cfg, hs, enc, emb)
        a_hs, a_enc = attention_block(sd, b + "attn1.", cfg, n_hs, n_enc, rope)
        hs = hs + gate * a_hs
        enc = enc + enc_gate * a_enc
        n_hs, n_enc, gate_ff, enc_gate_ff = layer_norm_zero(sd, b + "norm2.", cfg, hs, enc, emb)
        ff_in = torch.cat([n_enc, n_hs], dim=1)
        ff = F.linear(ff_in, sd[b + "ff.net.0.proj.weight"], sd[b + "ff.net.0.proj.bias"])
        ff = F.gelu(ff, approximate="tanh")
        ff = F.linear(ff, sd[b + "ff.net.2.weight"], sd[b + "ff.net.2.bias"])
        hs = hs + gate_ff * ff[:, St:]
        enc = enc + enc_gate_ff * ff[:, :St]
    x = torch.cat([enc, hs], dim=1)
    x = F.layer_norm(x, (D,), sd["norm_final.weight"], sd["norm_final.bias"], cfg.norm_eps)
    hs = x[:, St:]
    mod = F.linear(F.silu(emb), sd["norm_out.linear.weight"], sd["norm_out.linear.bias"])
    shift, scale = mod.chunk(2, dim=1)
    hs = F.layer_norm(hs, (D,), sd["norm_out.norm.weight"], sd["norm_out.norm.bias"], cfg.norm_eps) * (1 + scale)[:, None, :] + shift[:, None, :]
    hs = F.linear(hs, sd["proj_out.weight"], sd["proj_out.bias"])
    if pt is None:
        out = hs.reshape(B, Fr, H // p, W // p, -1, p, p).permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
    else:
        out = hs.reshape(B, (Fr + pt - 1) // pt, H // p, W // p, -1, pt, p, p)
        out = out.permute(0, 1, 5, 4, 2, 6, 3, 7).flatten(6, 7).flatten(4, 5).flatten(1, 2)
    return out


def block_forward(sd, cfg, i, hs, enc, emb, rope=None):
    """One CogVideoXBlock (used by the bounded CPU baseline in bench.py)."""
    b = f"transformer_blocks.{i}."
    St = enc.shape[1]
    n_hs, n_enc, gate, enc_gate = layer_norm_zero(sd, b + "norm1.", cfg, hs, enc, emb)
    a_hs, a_enc = attention_block(sd, b + "attn1.", cfg, n_hs, n_enc, rope)
    hs = hs + gate * a_hs
    enc = enc + enc_gate * a_enc
    n_hs, n_enc, gate_ff, enc_gate_ff = layer_norm_zero(sd, b + "norm2.", cfg, hs, enc, emb)
    ff = F.linear(torch.cat([n_enc, n_hs], dim=1), sd[b + "ff.net.0.proj.weight"], sd[b + "ff.net.0.proj.bias"])
    ff = F.linear(F.gelu(ff, approximate="tanh"), sd[b + "ff.net.2.weight"], sd[b + "ff.net.2.bias"])
    return hs + gate_ff * ff[:, St:], enc + enc_gate_ff * ff[:, :St]


# ------------------------------------------------------------------ App. A.4: schedulers
def cogvideox_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, snr_shift_scale=1.0) -> np.ndarray:
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float64) ** 2
    ac = np.cumprod(1.0 - betas)
    ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
    r = np.sqrt(ac)                                                      # rescale_betas_zero_snr
    r0, rT = r[0], r[-1]
    r = (r - rT) * r0 / (r0 - rT)
    return r ** 2


def trailing_timesteps(num_inference_steps: int, num_train_timesteps: int = 1000) -> np.ndarray:
    return (np.round(np.arange(num_train_timesteps, 0, -num_train_timesteps / num_inference_steps)) - 1).astype(np.int64)


def ddim_step(ac: np.ndarray, t: int, t_prev: int, sample: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """CogVideoXDDIMScheduler.step, v-prediction, eta = 0 (App. A.4). `sample` keeps its dtype for coef*sample."""
    a_t = float(ac[t])
    a_prev = float(ac[t_prev]) if t_prev >= 0 else 1.0
    x0 = (a_t ** 0.5) * sample - ((1 - a_t) ** 0.5) * v
    a = ((1 - a_prev) / (1 - a_t)) ** 0.5
    b = a_prev ** 0.5 - a_t ** 0.5 * a
    return a * sample + b * x0


def add_noise(ac, x, noise, t):
    sa = torch.as_tensor(ac[t] ** 0.5, dtype=x.dtype).reshape(-1, *[1] * (x.dim() - 1))
    sb = torch.as_tensor((1 - ac[t]) ** 0.5, dtype=x.dtype).reshape(-1, *[1] * (x.dim() - 1))
    return sa * x + sb * noise


def get_velocity(ac, x, noise, t):
    sa = torch.as_tensor(ac[t] ** 0.5, dtype=x.dtype).reshape(-1, *[1] * (x.dim() - 1))
    sb = torch.as_tensor((1 - ac[t]) ** 0.5, dtype=x.dtype).reshape(-1, *[1] * (x.dim() - 1))
    return sa * noise - sb * x


def dpm_coefficients(ac: np.ndarray, t: int, t_prev: int, t_back: int | None):
    """CogVideoXDPMScheduler multipliers (App. A.4): returns (m1, m2, m_noise, r or None)."""
    a_t = float(ac[t])
    a_prev = float(ac[t_prev]) if t_prev >= 0 else 1.0

    def lam(a):  # log-SNR/2 with torch's log(0) = -inf (zero-terminal SNR: ac[999] == 0) and log(inf) = +inf
        a = float(a)
        return -math.inf if a <= 0.0 else (math.inf if a >= 1.0 else math.log((a / (1 - a)) ** 0.5))

    lam_t = lam(a_t)
    h = lam(a_prev) - lam_t                                            # +inf on the first (a_t = 0) and last (a_prev = 1) step
    m1 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * math.exp(-h)
    m2 = math.expm1(-2 * h) * a_prev ** 0.5
    mn = (1 - a_prev) ** 0.5 * (1 - math.exp(-2 * h)) ** 0.5
    r = None
    if t_back is not None and not math.isinf(h):
        r = (lam_t - lam(ac[t_back])) / h                              # +inf when t_back = 999: 1/(2r) = 0
    return m1, m2, mn, r


def cfg_combine(noise_pred: torch.Tensor, guidance_scale: float) -> torch.Tensor:
    u, c = noise_pred.float().chunk(2)
    return u + guidance_scale * (c - u)


# ------------------------------------------------------------------ a-8: PEFT LoRA merge
def lora_merge(weight: torch.Tensor, A: torch.Tensor, B: torch.Tensor, scaling: float) -> torch.Tensor:
    """W' = W + scaling * (B @ A), fp32 product, one rounding to the weight dtype (generate/CogVideoX-5B.py:29-30)."""
    return (weight.float() + scaling * (B.float() @ A.float())).to(weight.dtype)
