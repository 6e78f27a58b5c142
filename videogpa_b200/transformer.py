"""CogVideoX 3-D transformer denoiser on the sm_100a kernels — the drop-in for
diffusers' CogVideoXTransformer3DModel as the reference uses it.

Reference call sites this mirrors (same keyword names, same return shape):
    pipe.transformer(hidden_states=..., encoder_hidden_states=..., timestep=..., image_rotary_emb=...,
                     return_dict=False)[0]                         diffusers pipeline, generate/CogVideoX-5B.py:72-77
    self.transformer(x, encoder_hidden_states=..., timestep=..., return_dict=True).sample
                                                                   train/CogVideoX-5B/03_train.py:134-151
The math follows SURVEY.md App. A.1/A.2. Every arithmetic step is one of the C-ABI kernels in
include/videogpa_b200.h (videogpa_b200/dense.py); torch is used for device memory and views only.
Per block (9 launches): adaLN GEMV -> LN+modulate -> fused QKV GEMM (bias + per-head LayerNorm + RoPE
epilogue) -> attention -> out-proj GEMM (gate * y + residual epilogue) -> adaLN GEMV -> LN+modulate
-> FF1 GEMM (bias + GELU-tanh epilogue) -> FF2 GEMM (gate * y + residual epilogue).
The residual stream is one [B, text+video, D] bf16 buffer (text rows first, as in the reference's
`cat([encoder_hidden_states, hidden_states])`), updated in place.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass
from types import SimpleNamespace

import torch

from . import dense

BF16 = torch.bfloat16


@dataclass
class TransformerConfig:
    num_attention_heads: int = 48
    attention_head_dim: int = 64
    in_channels: int = 16
    out_channels: int = 16
    time_embed_dim: int = 512
    text_embed_dim: int = 4096
    num_layers: int = 42
    patch_size: int = 2
    patch_size_t: int | None = None      # CogVideoX1.5: 2 (Linear patch embed over (c, pt, ph, pw))
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

    @classmethod
    def cogvideox_5b(cls) -> "TransformerConfig":
        return cls()

    @classmethod
    def cogvideox_5b_i2v(cls) -> "TransformerConfig":
        return cls(in_channels=32, use_learned_positional_embeddings=True)

    @classmethod
    def cogvideox1_5_5b(cls) -> "TransformerConfig":
        """CogVideoX1.5-5B T2V (generate/CogVideoX1.5-5B.py: 81 frames 1360x768): temporal patching, 224 text tokens."""
        return cls(patch_size_t=2, sample_width=170, sample_height=96, sample_frames=81, max_text_seq_length=224)


class Transformer3DOutput(SimpleNamespace):
    """`.sample` like diffusers' Transformer2DModelOutput."""


class _Block:
    __slots__ = ("n1_w", "n1_b", "n1_lw", "n1_lb", "n2_w", "n2_b", "n2_lw", "n2_lb", "w_qkv", "b_qkv", "lnq", "lnk",
                 "w_o", "b_o", "w_ff1", "b_ff1", "w_ff2", "b_ff2")


class CogVideoXTransformer3D:
    """Inference-only mirror of CogVideoXTransformer3DModel (weights frozen, bf16)."""

    def __init__(self, config: TransformerConfig, state_dict: dict, device="cuda"):
        if config.attention_head_dim != 64:
            raise RuntimeError("the sm_100a attention kernel is built for head_dim 64 (CogVideoX)")
        self.config = config
        self.device = torch.device(device)
        self.dtype = BF16
        self.training = False
        self._load(state_dict)

    # ------------------------------------------------------------------ construction
    @classmethod
    def random_init(cls, config: TransformerConfig, seed: int = 1234, device="cuda", std: float = 0.02) -> "CogVideoXTransformer3D":
        """SURVEY.md §8d synthetic weights (no checkpoints are reachable): N(0, std^2) Linear/Conv weights,
        zero biases, LayerNorm gamma 1 beta 0, drawn on the device."""
        dev = torch.device(device)
        g = torch.Generator(device=dev).manual_seed(seed)
        D, Tm, p = config.inner_dim, config.time_embed_dim, config.patch_size
        sd = {}

        def lin(name, o, i):
            sd[name + ".weight"] = (torch.randn(o, i, generator=g, device=dev, dtype=torch.float32) * std).to(BF16)
            sd[name + ".bias"] = torch.zeros(o, device=dev, dtype=BF16)

        def norm(name, n):
            sd[name + ".weight"] = torch.ones(n, device=dev, dtype=BF16)
            sd[name + ".bias"] = torch.zeros(n, device=dev, dtype=BF16)

        if config.patch_size_t is None:
            sd["patch_embed.proj.weight"] = (torch.randn(D, config.in_channels, p, p, generator=g, device=dev) * std).to(BF16)
        else:
            sd["patch_embed.proj.weight"] = (torch.randn(D, config.in_channels * config.patch_size_t * p * p, generator=g, device=dev) * std).to(BF16)
        sd["patch_embed.proj.bias"] = torch.zeros(D, device=dev, dtype=BF16)
        lin("patch_embed.text_proj", D, config.text_embed_dim)
        if config.use_learned_positional_embeddings:
            n_tok = config.max_text_seq_length + ((config.sample_frames - 1) // config.temporal_compression_ratio + 1) * \
                (config.sample_height // p) * (config.sample_width // p)
            sd["patch_embed.pos_embedding"] = (torch.randn(1, n_tok, D, generator=g, device=dev) * std).to(BF16)
        lin("time_embedding.linear_1", Tm, D)
        lin("time_embedding.linear_2", Tm, Tm)
        for i in range(config.num_layers):
            b = f"transformer_blocks.{i}."
            lin(b + "norm1.linear", 6 * D, Tm); norm(b + "norm1.norm", D)
            lin(b + "norm2.linear", 6 * D, Tm); norm(b + "norm2.norm", D)
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(b + "attn1." + n, D, D)
            norm(b + "attn1.norm_q", 64); norm(b + "attn1.norm_k", 64)
            lin(b + "ff.net.0.proj", config.ffn_mult * D, D)
            lin(b + "ff.net.2", D, config.ffn_mult * D)
        norm("norm_final", D)
        lin("norm_out.linear", 2 * D, Tm); norm("norm_out.norm", D)
        lin("proj_out", p * p * (config.patch_size_t or 1) * config.out_channels, D)
        return cls(config, sd, device=dev)

    def _load(self, sd: dict) -> None:
        dev, c = self.device, self.config
        D = c.inner_dim

        def w(name):
            if name not in sd:
                raise RuntimeError(f"state dict is missing {name}")
            return sd[name].to(device=dev, dtype=BF16).contiguous()

        def f32(name):
            return sd[name].to(device=dev, dtype=torch.float32).contiguous()

        self.patch_w = w("patch_embed.proj.weight").reshape(D, -1).contiguous()          # [D, C*p*p], (c, ph, pw) order
        self.patch_b = w("patch_embed.proj.bias") if "patch_embed.proj.bias" in sd else torch.zeros(D, device=dev, dtype=BF16)
        pt = c.patch_size_t
        if pt is not None:
            # 1.5: features are (c, pt, ph, pw); the per-frame patchify kernel produces (c, ph, pw) per frame, so the two
            # frames of a temporal patch are concatenated as (pt, c, ph, pw) and the weight columns are permuted to match
            pp = c.patch_size ** 2
            self.patch_w = self.patch_w.view(D, c.in_channels, pt, pp).permute(0, 2, 1, 3).reshape(D, -1).contiguous()
        self.text_w, self.text_b = w("patch_embed.text_proj.weight"), w("patch_embed.text_proj.bias")
        self.pos_embedding = w("patch_embed.pos_embedding") if c.use_learned_positional_embeddings else None
        self.t1_w, self.t1_b = w("time_embedding.linear_1.weight"), w("time_embedding.linear_1.bias")
        self.t2_w, self.t2_b = w("time_embedding.linear_2.weight"), w("time_embedding.linear_2.bias")
        self.blocks: list[_Block] = []
        for i in range(c.num_layers):
            p = f"transformer_blocks.{i}."
            b = _Block()
            b.n1_lw, b.n1_lb = w(p + "norm1.linear.weight"), w(p + "norm1.linear.bias")
            b.n1_w, b.n1_b = w(p + "norm1.norm.weight"), w(p + "norm1.norm.bias")
            b.n2_lw, b.n2_lb = w(p + "norm2.linear.weight"), w(p + "norm2.linear.bias")
            b.n2_w, b.n2_b = w(p + "norm2.norm.weight"), w(p + "norm2.norm.bias")
            b.w_qkv = torch.cat([w(p + "attn1.to_q.weight"), w(p + "attn1.to_k.weight"), w(p + "attn1.to_v.weight")], 0).contiguous()
            b.b_qkv = torch.cat([w(p + "attn1.to_q.bias"), w(p + "attn1.to_k.bias"), w(p + "attn1.to_v.bias")], 0).contiguous()
            b.lnq = (f32(p + "attn1.norm_q.weight"), f32(p + "attn1.norm_q.bias"))
            b.lnk = (f32(p + "attn1.norm_k.weight"), f32(p + "attn1.norm_k.bias"))
            b.w_o, b.b_o = w(p + "attn1.to_out.0.weight"), w(p + "attn1.to_out.0.bias")
            b.w_ff1, b.b_ff1 = w(p + "ff.net.0.proj.weight"), w(p + "ff.net.0.proj.bias")
            b.w_ff2, b.b_ff2 = w(p + "ff.net.2.weight"), w(p + "ff.net.2.bias")
            self.blocks.append(b)
        self.nf_w, self.nf_b = w("norm_final.weight"), w("norm_final.bias")
        self.no_lw, self.no_lb = w("norm_out.linear.weight"), w("norm_out.linear.bias")
        self.no_w, self.no_b = w("norm_out.norm.weight"), w("norm_out.norm.bias")
        self.po_w, self.po_b = w("proj_out.weight"), w("proj_out.bias")
        if pt is not None:
            pp, Co = c.patch_size ** 2, c.out_channels
            self.po_w = self.po_w.view(Co, pt, pp, D).permute(1, 0, 2, 3).reshape(-1, D).contiguous()
            self.po_b = self.po_b.view(Co, pt, pp).permute(1, 0, 2).reshape(-1).contiguous()

    # ------------------------------------------------------------------ nn.Module-ish surface the callers touch
    def eval(self):
        self.training = False
        return self

    def requires_grad_(self, flag: bool = False):
        if flag:
            raise RuntimeError("videogpa_b200 transformer is forward-only in this round (no backward kernels)")
        return self

    def enable_gradient_checkpointing(self):
        return None

    def to(self, *a, **k):
        return self

    def attention_weight(self, layer: int, module: str) -> torch.Tensor:
        """Writable [D, D] view of attn1.{to_q,to_k,to_v,to_out.0}.weight (used by the LoRA merge)."""
        b, D = self.blocks[layer], self.config.inner_dim
        if module == "to_out.0":
            return b.w_o
        idx = {"to_q": 0, "to_k": 1, "to_v": 2}[module]
        return b.w_qkv[idx * D:(idx + 1) * D]

    # ------------------------------------------------------------------ forward
    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, timestep,
                timestep_cond=None, image_rotary_emb=None, attention_kwargs=None, return_dict: bool = True,
                num_layers: int | None = None):
        c = self.config
        if timestep_cond is not None:
            raise RuntimeError("timestep_cond is not used by CogVideoX checkpoints and is not supported")
        if hidden_states.dim() != 5:
            raise RuntimeError("hidden_states must be [B, F, C, H, W]")
        B, Fr, C, H, W = hidden_states.shape
        if C != c.in_channels:
            raise RuntimeError(f"hidden_states has {C} channels, config.in_channels = {c.in_channels}")
        p, D, heads = c.patch_size, c.inner_dim, c.num_attention_heads
        dev = self.device
        hs = hidden_states.to(device=dev, dtype=BF16).contiguous()
        enc_in = encoder_hidden_states.to(device=dev, dtype=BF16).contiguous()
        St = enc_in.shape[1]
        pt = c.patch_size_t or 1
        if Fr % pt != 0:
            raise RuntimeError(f"the number of latent frames ({Fr}) must be a multiple of patch_size_t ({pt}); the pipeline pads it")
        hw = (H // p) * (W // p)
        Sv = (Fr // pt) * hw
        S = St + Sv
        ts = torch.as_tensor(timestep, device=dev).reshape(-1).to(torch.float32)
        if ts.numel() == 1 and B > 1:
            ts = ts.expand(B).contiguous()

        # time embedding: sinusoid -> linear_1 -> SiLU -> linear_2                    (App. A.1)
        t_emb = dense.timestep_embedding(ts, D)
        e1 = dense.linear_smallm(t_emb, self.t1_w, self.t1_b)
        emb = dense.linear_smallm(e1, self.t2_w, self.t2_b, act_in=dense.ACT_SILU)

        # patch embed: [text_proj(enc) ; proj(2x2 patches)] (+ learned positional embedding for I2V)
        x = torch.empty((B, S, D), dtype=BF16, device=dev)
        patches = dense.patchify(hs.view(B * Fr, C, H, W))                                   # [B*Fr*hw, C*p*p]
        if pt > 1:                                                                            # [B, Fr/pt, pt, hw, Cpp] -> [.., hw, pt*Cpp]
            patches = patches.view(B, Fr // pt, pt, hw, -1).permute(0, 1, 3, 2, 4).reshape(B * Sv, -1).contiguous()
        epi = dense.EPI_BIAS
        if self.pos_embedding is not None:
            if self.pos_embedding.shape[1] < S:
                raise RuntimeError("pos_embedding is shorter than the token sequence")
            x.copy_(self.pos_embedding[:, :S].expand(B, S, D))
            epi = dense.EPI_GATE_RES                       # x <- pos + y (gate pointers NULL = 1)
        for b in range(B):
            dense.linear(enc_in[b], self.text_w, self.text_b, out=x[b, :St], epilogue=epi)
            dense.linear(patches[b * Sv:(b + 1) * Sv], self.patch_w, self.patch_b, out=x[b, St:], epilogue=epi)

        rope = None
        if image_rotary_emb is not None:
            cos, sin = image_rotary_emb
            rope = (cos.to(device=dev, dtype=torch.float32).contiguous(), sin.to(device=dev, dtype=torch.float32).contiguous())
            if tuple(rope[0].shape) != (Sv, 64):
                raise RuntimeError(f"image_rotary_emb must be ([{Sv}, 64], [{Sv}, 64]), got {tuple(rope[0].shape)}")

        x2 = x.view(B * S, D)
        n = torch.empty_like(x2)
        qkv = torch.empty((B, S, 3 * D), dtype=BF16, device=dev)
        att = torch.empty((B, S, D), dtype=BF16, device=dev)
        ffh = torch.empty((B * S, c.ffn_mult * D), dtype=BF16, device=dev)
        seg = dict(rows_per_sample=S, text_rows=St)
        L = c.num_layers if num_layers is None else num_layers
        for blk in self.blocks[:L]:
            # norm1: shift, scale, gate, enc_shift, enc_scale, enc_gate = linear(SiLU(emb)).chunk(6)
            m = dense.linear_smallm(emb, blk.n1_lw, blk.n1_lb, act_in=dense.ACT_SILU)
            dense.layernorm_modulate(x2, blk.n1_w, blk.n1_b, eps=c.norm_eps, out=n, **seg,
                                     shift_vid=m[:, 0:D], scale_vid=m[:, D:2 * D], shift_txt=m[:, 3 * D:4 * D],
                                     scale_txt=m[:, 4 * D:5 * D], mod_stride_b=6 * D)
            dense.linear(n, blk.w_qkv, blk.b_qkv, out=qkv.view(B * S, 3 * D), epilogue=dense.EPI_QKV, **seg,
                         ln_q=blk.lnq, ln_k=blk.lnk, ln_eps=1e-6, rope=rope, model_dim=D)
            dense.attention(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], heads, out=att)
            dense.linear(att.view(B * S, D), blk.w_o, blk.b_o, out=x2, epilogue=dense.EPI_GATE_RES, **seg,
                         gate_vid=m[:, 2 * D:3 * D], gate_txt=m[:, 5 * D:6 * D], gate_stride_b=6 * D)
            # norm2 + feed-forward
            m2 = dense.linear_smallm(emb, blk.n2_lw, blk.n2_lb, act_in=dense.ACT_SILU)
            dense.layernorm_modulate(x2, blk.n2_w, blk.n2_b, eps=c.norm_eps, out=n, **seg,
                                     shift_vid=m2[:, 0:D], scale_vid=m2[:, D:2 * D], shift_txt=m2[:, 3 * D:4 * D],
                                     scale_txt=m2[:, 4 * D:5 * D], mod_stride_b=6 * D)
            dense.linear(n, blk.w_ff1, blk.b_ff1, out=ffh, epilogue=dense.EPI_BIAS_GELU)
            dense.linear(ffh, blk.w_ff2, blk.b_ff2, out=x2, epilogue=dense.EPI_GATE_RES, **seg,
                         gate_vid=m2[:, 2 * D:3 * D], gate_txt=m2[:, 5 * D:6 * D], gate_stride_b=6 * D)

        # norm_final over [text; video], AdaLayerNorm (norm_out) and proj_out on the video rows
        dense.layernorm_modulate(x2, self.nf_w, self.nf_b, eps=c.norm_eps, out=n)
        mo = dense.linear_smallm(emb, self.no_lw, self.no_lb, act_in=dense.ACT_SILU)      # shift, scale
        y = dense.layernorm_modulate(n, self.no_w, self.no_b, eps=c.norm_eps, out=x2, **seg,
                                     shift_vid=mo[:, 0:D], scale_vid=mo[:, D:2 * D], shift_txt=mo[:, 0:D],
                                     scale_txt=mo[:, D:2 * D], mod_stride_b=2 * D)
        y3 = y.view(B, S, D)
        pd = p * p * c.out_channels
        tok = torch.empty((B * Sv, pt * pd), dtype=BF16, device=dev)
        for b in range(B):
            dense.linear(y3[b, St:], self.po_w, self.po_b, out=tok[b * Sv:(b + 1) * Sv])
        if pt > 1:                                                                            # [.., hw, pt, pd] -> one row per (frame, patch)
            tok = tok.view(B, Fr // pt, hw, pt, pd).permute(0, 1, 3, 2, 4).reshape(B * Fr * hw, pd).contiguous()
        out = dense.unpatchify(tok, B * Fr, c.out_channels, H, W).view(B, Fr, c.out_channels, H, W)
        if not return_dict:
            return (out,)
        return Transformer3DOutput(sample=out)

    __call__ = forward

    # ------------------------------------------------------------------ bookkeeping for bench.py
    def flops_per_sample(self, St: int, Sv: int, num_layers: int | None = None) -> float:
        """Algorithmic FLOPs of one sample-forward."""
        c = self.config
        D, S = c.inner_dim, St + Sv
        L = c.num_layers if num_layers is None else num_layers
        lin = 2.0 * S * D * (3 * D + D + 2 * c.ffn_mult * D)
        attn = 4.0 * S * S * D
        embed = 2.0 * Sv * D * (c.in_channels * c.patch_size ** 2) + 2.0 * St * D * c.text_embed_dim + \
            2.0 * Sv * D * (c.patch_size ** 2 * c.out_channels)
        return L * (lin + attn) + embed

    def kernel_launches(self, B: int, num_layers: int | None = None) -> int:
        """Kernels of this library launched by one forward: per block 2 modulation linears, 2 LayerNorm-modulate, 4 GEMMs and the
        attention call's 3 kernels (|q|,|k| bound pre-pass, bounded-softmax kernel, exact kernel for the heads above the bound)."""
        L = self.config.num_layers if num_layers is None else num_layers
        return 3 + 1 + 2 * B + 11 * L + 3 + B + 1

    def config_dict(self) -> dict:
        return asdict(self.config)
