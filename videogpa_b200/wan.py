"""Wan2.2 TI2V-5B diffusion transformer on the sm_100a kernels — the drop-in for `WanModel.forward` as the reference
drives it (generate/Wan2.2-TI2V-5B.py:120-129 through WanTI2V.generate; train/Wan2.2-TI2V-5B/03_train.py:228-233;
model dims train/Wan2.2-TI2V-5B/03_train.py:9-13).

Same call shape as the Wan repo: `model(x=[latent [48,F,H,W]], t=..., context=[text [L,4096]], seq_len=...) -> [tensor]`,
one sample per forward (Wan evaluates the cond and uncond branches as separate forwards). Every arithmetic step is a
C-ABI kernel: the q|k|v, o, cross-attention and FFN linears run on the tcgen05 GEMM (bias / GELU-tanh / gated-residual
epilogues), self- and cross-attention on the head_dim-128 attention kernel, LayerNorm + modulation, RMSNorm(q/k) + RoPE
and the modulation sums on HBM-bound row kernels.

Per-token timesteps (Wan2.2 TI2V: the first latent frame is the conditioning image and carries t = 0,
train/Wan2.2-TI2V-5B/03_train.py:119-125) take only two values, so modulation is computed for two row segments
(first-frame tokens | the rest) and selected per row inside the kernels.

Precision plan = the Wan repo's (torch.autocast(bf16) around an fp32 model input): the residual stream x is fp32, every
linear consumes and produces bf16 with fp32 accumulation, LayerNorm + modulation are evaluated in fp32 on the fp32 stream
and rounded to bf16 once for the next linear, and the gated residual updates add in fp32 (vgpa_linear_bf16 with
VGPA_EPI_GATE_RES_F32, vgpa_layernorm_modulate_bf16 with x_is_f32). tests/test_gpu_parity_full.py measures the full-size
forward against the fp32 oracle next to the same oracle under autocast.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import _lib, dense

BF16 = torch.bfloat16


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
    def head_dim(self) -> int:
        return self.dim // self.num_heads

    @classmethod
    def ti2v_5b(cls) -> "WanConfig":
        return cls()


def rope_tables(cfg: WanConfig, f: int, h: int, w: int, device="cpu"):
    """(cos, sin) [f*h*w, head_dim] fp32, repeat-interleaved over the complex pairs: head_dim split d-4(d//6) | 2(d//6) | 2(d//6)
    over (t, h, w), theta 10000 (rope_params / rope_apply of the Wan repo, App. A.7)."""
    d = cfg.head_dim
    dims = [d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)]

    def ang(n, dim):
        return torch.outer(torch.arange(n, dtype=torch.float64), 1.0 / torch.pow(10000, torch.arange(0, dim, 2, dtype=torch.float64) / dim))

    at, ah, aw = ang(f, dims[0]), ang(h, dims[1]), ang(w, dims[2])
    a = torch.cat([at[:, None, None, :].expand(f, h, w, -1), ah[None, :, None, :].expand(f, h, w, -1),
                   aw[None, None, :, :].expand(f, h, w, -1)], dim=-1).reshape(f * h * w, d // 2)
    return (a.cos().repeat_interleave(2, dim=1).float().contiguous().to(device),
            a.sin().repeat_interleave(2, dim=1).float().contiguous().to(device))


def _rmsnorm_rope(x: torch.Tensor, weight: torch.Tensor, eps: float, rope=None, head_dim: int = 128) -> None:
    """In place on a [rows, D] bf16 view (row stride may exceed D)."""
    lib = _lib.load()
    if x.dtype != BF16 or not x.is_cuda or x.stride(1) != 1:
        raise RuntimeError("rmsnorm_rope needs a CUDA bf16 view with a contiguous last dimension")
    cos = rope[0].data_ptr() if rope is not None else None
    sin = rope[1].data_ptr() if rope is not None else None
    _lib.check(lib.vgpa_rmsnorm_rope_bf16(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), weight.data_ptr(), eps, cos, sin,
                                          head_dim, 0, _lib.current_stream()), "vgpa_rmsnorm_rope_bf16")


def _add_rows(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16(a[r, :] + b[:]) — a [R, N] bf16, b [N] fp32."""
    lib = _lib.load()
    out = torch.empty(a.shape, dtype=BF16, device=a.device)
    _lib.check(lib.vgpa_add_rows_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.shape[0], a.shape[1], a.stride(0),
                                      _lib.current_stream()), "vgpa_add_rows_bf16")
    return out


class _Block:
    __slots__ = ("w_qkv", "b_qkv", "nq", "nk", "w_o", "b_o", "n3_w", "n3_b", "w_cq", "b_cq", "w_ckv", "b_ckv", "cnq", "cnk",
                 "w_co", "b_co", "w_f0", "b_f0", "w_f2", "b_f2", "mod")


class WanTransformer3D:
    """Inference-only mirror of WanModel (weights frozen, bf16)."""

    def __init__(self, config: WanConfig, state_dict: dict, device="cuda"):
        if config.head_dim != 128:
            raise RuntimeError("the Wan path is built for head_dim 128")
        if tuple(config.patch_size) != (1, 2, 2):
            raise RuntimeError("only patch_size (1, 2, 2) is supported")
        self.config = config
        self.device = torch.device(device)
        self.dtype = BF16
        self._load(state_dict)

    @classmethod
    def random_init(cls, config: WanConfig, seed: int = 21, device="cuda", std: float = 0.02) -> "WanTransformer3D":
        dev = torch.device(device)
        g = torch.Generator(device=dev).manual_seed(seed)
        sd, D = {}, config.dim

        def lin(name, o, i):
            sd[name + ".weight"] = (torch.randn(o, i, generator=g, device=dev) * std).to(BF16)
            sd[name + ".bias"] = torch.zeros(o, device=dev, dtype=BF16)

        sd["patch_embedding.weight"] = (torch.randn(D, config.in_dim, 1, 2, 2, generator=g, device=dev) * std).to(BF16)
        sd["patch_embedding.bias"] = torch.zeros(D, device=dev, dtype=BF16)
        lin("text_embedding.0", D, config.text_dim); lin("text_embedding.2", D, D)
        lin("time_embedding.0", D, config.freq_dim); lin("time_embedding.2", D, D)
        lin("time_projection.1", 6 * D, D)
        for i in range(config.num_layers):
            b = f"blocks.{i}."
            for a in ("self_attn", "cross_attn"):
                for m in ("q", "k", "v", "o"):
                    lin(b + f"{a}.{m}", D, D)
                sd[b + f"{a}.norm_q.weight"] = torch.ones(D, device=dev)
                sd[b + f"{a}.norm_k.weight"] = torch.ones(D, device=dev)
            sd[b + "norm3.weight"] = torch.ones(D, device=dev, dtype=BF16)
            sd[b + "norm3.bias"] = torch.zeros(D, device=dev, dtype=BF16)
            lin(b + "ffn.0", config.ffn_dim, D); lin(b + "ffn.2", D, config.ffn_dim)
            sd[b + "modulation"] = torch.randn(1, 6, D, generator=g, device=dev) / D ** 0.5
        lin("head.head", config.out_dim * 4, D)
        sd["head.modulation"] = torch.randn(1, 2, D, generator=g, device=dev) / D ** 0.5
        return cls(config, sd, device=dev)

    def _load(self, sd: dict) -> None:
        dev, c = self.device, self.config
        D = c.dim

        def w(name):
            if name not in sd:
                raise RuntimeError(f"state dict is missing {name}")
            return sd[name].to(device=dev, dtype=BF16).contiguous()

        def f32(name):
            return sd[name].to(device=dev, dtype=torch.float32).contiguous()

        self.patch_w = w("patch_embedding.weight").reshape(D, -1).contiguous()           # [D, C*1*2*2], (c, ph, pw) order
        self.patch_b = w("patch_embedding.bias")
        self.te0_w, self.te0_b = w("text_embedding.0.weight"), w("text_embedding.0.bias")
        self.te2_w, self.te2_b = w("text_embedding.2.weight"), w("text_embedding.2.bias")
        self.ti0_w, self.ti0_b = w("time_embedding.0.weight"), w("time_embedding.0.bias")
        self.ti2_w, self.ti2_b = w("time_embedding.2.weight"), w("time_embedding.2.bias")
        self.tp_w, self.tp_b = w("time_projection.1.weight"), w("time_projection.1.bias")
        self.blocks: list[_Block] = []
        for i in range(c.num_layers):
            p = f"blocks.{i}."
            b = _Block()
            b.w_qkv = torch.cat([w(p + "self_attn.q.weight"), w(p + "self_attn.k.weight"), w(p + "self_attn.v.weight")], 0).contiguous()
            b.b_qkv = torch.cat([w(p + "self_attn.q.bias"), w(p + "self_attn.k.bias"), w(p + "self_attn.v.bias")], 0).contiguous()
            b.nq, b.nk = f32(p + "self_attn.norm_q.weight"), f32(p + "self_attn.norm_k.weight")
            b.w_o, b.b_o = w(p + "self_attn.o.weight"), w(p + "self_attn.o.bias")
            b.n3_w, b.n3_b = w(p + "norm3.weight"), w(p + "norm3.bias")
            b.w_cq, b.b_cq = w(p + "cross_attn.q.weight"), w(p + "cross_attn.q.bias")
            b.w_ckv = torch.cat([w(p + "cross_attn.k.weight"), w(p + "cross_attn.v.weight")], 0).contiguous()
            b.b_ckv = torch.cat([w(p + "cross_attn.k.bias"), w(p + "cross_attn.v.bias")], 0).contiguous()
            b.cnq, b.cnk = f32(p + "cross_attn.norm_q.weight"), f32(p + "cross_attn.norm_k.weight")
            b.w_co, b.b_co = w(p + "cross_attn.o.weight"), w(p + "cross_attn.o.bias")
            b.w_f0, b.b_f0 = w(p + "ffn.0.weight"), w(p + "ffn.0.bias")
            b.w_f2, b.b_f2 = w(p + "ffn.2.weight"), w(p + "ffn.2.bias")
            b.mod = f32(p + "modulation").reshape(6 * D)
            self.blocks.append(b)
        # head: rows permuted so the output features come out as (c, ph, pw) — the order the unpatchify kernel reads —
        # instead of the Wan repo's (pt, ph, pw, c)
        hw_, hb_ = w("head.head.weight"), w("head.head.bias")
        C_ = c.out_dim
        perm = torch.tensor([(q * 2 + r) * C_ + ch for ch in range(C_) for q in range(2) for r in range(2)], device=dev)
        pad = (-perm.numel()) % 64
        self.head_w = torch.cat([hw_[perm], torch.zeros(pad, D, device=dev, dtype=BF16)], 0).contiguous()
        self.head_b = torch.cat([hb_[perm], torch.zeros(pad, device=dev, dtype=BF16)], 0).contiguous()
        self.head_mod = f32("head.modulation").reshape(2 * D)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def attention_weight(self, layer: int, module: str) -> torch.Tensor:
        """Writable [D, D] view for the LoRA merge: targets q, k, v, o of self_attn and cross_attn
        (train/Wan2.2-TI2V-5B/03_train.py:82)."""
        b, D = self.blocks[layer], self.config.dim
        kind, m = module.split(".")
        if kind == "self_attn":
            return b.w_o if m == "o" else b.w_qkv[{"q": 0, "k": 1, "v": 2}[m] * D:][:D]
        if kind == "cross_attn":
            if m == "q":
                return b.w_cq
            return b.w_co if m == "o" else b.w_ckv[{"k": 0, "v": 1}[m] * D:][:D]
        raise RuntimeError(f"unknown attention module {module}")

    # ------------------------------------------------------------------ forward (one sample)
    def _forward_one(self, lat: torch.Tensor, t_first: float, t_rest: float, context: torch.Tensor, num_layers=None) -> torch.Tensor:
        c = self.config
        dev, D, n_heads, hd = self.device, c.dim, c.num_heads, c.head_dim
        Cin, F_, H, W = lat.shape
        if Cin != c.in_dim or H % 2 or W % 2:
            raise RuntimeError(f"latent must be [{c.in_dim}, F, even H, even W], got {tuple(lat.shape)}")
        h, w_ = H // 2, W // 2
        hw = h * w_
        S = F_ * hw
        frames = lat.to(device=dev, dtype=BF16).permute(1, 0, 2, 3).contiguous()                      # [F, C, H, W]
        x = dense.linear(dense.patchify(frames), self.patch_w, self.patch_b).float()                   # [S, D] fp32 residual stream
        # time embedding for the two distinct timesteps: row 0 = first-frame tokens, row 1 = the rest
        ts = torch.tensor([t_first, t_rest], dtype=torch.float32, device=dev)
        sin_emb = dense.timestep_embedding(ts, c.freq_dim)
        e = dense.linear_smallm(dense.linear_smallm(sin_emb, self.ti0_w, self.ti0_b), self.ti2_w, self.ti2_b, act_in=dense.ACT_SILU)
        e0 = dense.linear_smallm(e, self.tp_w, self.tp_b, act_in=dense.ACT_SILU)                       # [2, 6D]
        # text embedding (zero-padded to text_len like the Wan repo)
        ctx_in = torch.zeros((c.text_len, c.text_dim), dtype=BF16, device=dev)
        L_ = min(context.shape[0], c.text_len)
        ctx_in[:L_] = context[:L_].to(device=dev, dtype=BF16)
        ctx = dense.linear(dense.linear(ctx_in, self.te0_w, self.te0_b, epilogue=dense.EPI_BIAS_GELU), self.te2_w, self.te2_b)
        rope = rope_tables(c, F_, h, w_, device=dev)

        n = torch.empty((S, D), dtype=BF16, device=dev)
        qkv = torch.empty((S, 3 * D), dtype=BF16, device=dev)
        att = torch.empty((1, S, D), dtype=BF16, device=dev)
        qc = torch.empty((S, D), dtype=BF16, device=dev)
        kvc = torch.empty((c.text_len, 2 * D), dtype=BF16, device=dev)
        ffh = torch.empty((S, c.ffn_dim), dtype=BF16, device=dev)
        seg = dict(rows_per_sample=S, text_rows=hw)
        L = c.num_layers if num_layers is None else num_layers
        for blk in self.blocks[:L]:
            m = _add_rows(e0, blk.mod)                                                                 # [2, 6D]: rows = segments
            mod = lambda k, s_: m[s_:s_ + 1, k * D:(k + 1) * D]
            dense.layernorm_modulate(x, None, None, eps=c.eps, out=n, **seg, shift_txt=mod(0, 0), scale_txt=mod(1, 0),
                                     shift_vid=mod(0, 1), scale_vid=mod(1, 1), mod_stride_b=0)
            dense.linear(n, blk.w_qkv, blk.b_qkv, out=qkv)
            _rmsnorm_rope(qkv[:, :D], blk.nq, c.eps, rope, hd)
            _rmsnorm_rope(qkv[:, D:2 * D], blk.nk, c.eps, rope, hd)
            q3 = qkv.view(1, S, 3 * D)
            dense.attention(q3[..., :D], q3[..., D:2 * D], q3[..., 2 * D:], n_heads, out=att, head_dim=hd)
            dense.linear(att.view(S, D), blk.w_o, blk.b_o, out=x, epilogue=dense.EPI_GATE_RES_F32, **seg,
                         gate_txt=mod(2, 0), gate_vid=mod(2, 1), gate_stride_b=0)
            # cross-attention over the embedded text
            dense.layernorm_modulate(x, blk.n3_w, blk.n3_b, eps=c.eps, out=n)
            dense.linear(n, blk.w_cq, blk.b_cq, out=qc)
            _rmsnorm_rope(qc, blk.cnq, c.eps)
            dense.linear(ctx, blk.w_ckv, blk.b_ckv, out=kvc)
            _rmsnorm_rope(kvc[:, :D], blk.cnk, c.eps)
            k3 = kvc.view(1, c.text_len, 2 * D)
            dense.attention(qc.view(1, S, D), k3[..., :D], k3[..., D:], n_heads, out=att, head_dim=hd)
            dense.linear(att.view(S, D), blk.w_co, blk.b_co, out=x, epilogue=dense.EPI_GATE_RES_F32)
            # feed-forward
            dense.layernorm_modulate(x, None, None, eps=c.eps, out=n, **seg, shift_txt=mod(3, 0), scale_txt=mod(4, 0),
                                     shift_vid=mod(3, 1), scale_vid=mod(4, 1), mod_stride_b=0)
            dense.linear(n, blk.w_f0, blk.b_f0, out=ffh, epilogue=dense.EPI_BIAS_GELU)
            dense.linear(ffh, blk.w_f2, blk.b_f2, out=x, epilogue=dense.EPI_GATE_RES_F32, **seg,
                         gate_txt=mod(5, 0), gate_vid=mod(5, 1), gate_stride_b=0)
        # head: LN (1 + e1) + e0 with e = head.modulation + time embedding (before the projection)
        hm = _add_rows(torch.cat([e, e], dim=1), self.head_mod)                                        # [2, 2D]
        dense.layernorm_modulate(x, None, None, eps=c.eps, out=n, **seg, shift_txt=hm[0:1, :D], scale_txt=hm[0:1, D:],
                                 shift_vid=hm[1:2, :D], scale_vid=hm[1:2, D:], mod_stride_b=0)
        y = dense.linear(n, self.head_w, self.head_b)                                                  # [S, >= 4*C] in (c, ph, pw) order
        out = dense.unpatchify(y, F_, c.out_dim, H, W)                                                 # [F, C, H, W]
        return out.permute(1, 0, 2, 3).contiguous()

    @staticmethod
    def _two_timesteps(t, S: int, hw: int):
        """t: scalar / [1] / [S] per-token timesteps -> (t_first_frame, t_rest); per-token values must be constant inside
        the first frame and inside the rest (the Wan2.2 TI2V pattern)."""
        t = torch.as_tensor(t, dtype=torch.float32).reshape(-1).cpu()
        if t.numel() == 1:
            return float(t[0]), float(t[0])
        if t.numel() != S:
            raise RuntimeError(f"per-token timesteps must have {S} entries, got {t.numel()}")
        a, b = t[:hw], t[hw:]
        if (a != a[0]).any() or (b.numel() and (b != b[0]).any()):
            raise RuntimeError("per-token timesteps must be constant on the first latent frame and on the remaining frames")
        return float(a[0]), float(b[0]) if b.numel() else float(a[0])

    @torch.no_grad()
    def forward(self, x, t, context, seq_len=None, num_layers=None):
        """x: list of [C, F, H, W]; t: [B] or [B, seq_len]; context: list of [L, text_dim] -> list of [C, F, H, W] (bf16)."""
        if isinstance(x, torch.Tensor):
            x = list(x) if x.dim() == 5 else [x]
        if isinstance(context, torch.Tensor):
            context = list(context) if context.dim() == 3 else [context]
        if len(context) != len(x):
            raise RuntimeError("x and context must have the same batch size")
        t = torch.as_tensor(t)
        outs = []
        for i, (lat, ctx) in enumerate(zip(x, context)):
            S = lat.shape[1] * (lat.shape[2] // 2) * (lat.shape[3] // 2)
            if seq_len is not None and seq_len < S:
                raise RuntimeError(f"seq_len {seq_len} is smaller than the {S} tokens of the input")
            ti = t if t.dim() == 0 else t[i]
            if ti.dim() == 1 and ti.numel() > S:
                ti = ti[:S]
            tf, tr = self._two_timesteps(ti, S, (lat.shape[2] // 2) * (lat.shape[3] // 2))
            outs.append(self._forward_one(lat, tf, tr, ctx, num_layers))
        return outs

    __call__ = forward

    # ------------------------------------------------------------------ bookkeeping for bench.py
    def flops_per_forward(self, S: int, num_layers: int | None = None) -> float:
        c = self.config
        D, L = c.dim, (c.num_layers if num_layers is None else num_layers)
        lin = 2.0 * S * D * (3 * D + D + D + D + 2 * c.ffn_dim) + 2.0 * c.text_len * D * 2 * D
        attn = 4.0 * S * S * D + 4.0 * S * c.text_len * D
        return L * (lin + attn)


def flow_sigmas(num_steps: int, shift: float = 5.0, num_train_timesteps: int = 1000):
    """Shifted flow-matching sigma schedule (generate/Wan2.2-TI2V-5B.py:145 `--shift 5.0`)."""
    s = [1.0 + (1.0 / num_train_timesteps - 1.0) * k / max(1, num_steps - 1) for k in range(num_steps)]
    s = [shift * v / (1 + (shift - 1) * v) for v in s]
    return s + [0.0]


class WanDenoiseStep:
    """One guided denoise step of the Wan sampler loop: cond and uncond forwards (separate, as in WanTI2V.generate),
    `uncond + g (cond - uncond)` and a flow-matching Euler update, fused in vgpa_cfg_scheduler_step. `cfg_group`
    (parallel.CfgPairGroup) shards the two branches over two ranks with one all-gather of the prediction per step.
    `guided_velocity` + schedulers.FlowUniPCMultistepScheduler is the reference's default UniPC sampler (generate CLI)."""

    def __init__(self, model: WanTransformer3D, guide_scale: float = 5.0):
        self.model, self.guide_scale = model, guide_scale

    @torch.no_grad()
    def guided_velocity(self, latent, t_tokens, context, context_null) -> torch.Tensor:
        """cond and uncond forwards + `uncond + g (cond - uncond)` in fp32 (WanTI2V.generate's noise_pred): the input of a
        multistep sampler such as schedulers.FlowUniPCMultistepScheduler, which keeps the latent in fp32 between steps."""
        lat = latent.to(device=self.model.device, dtype=BF16).contiguous()
        cond = self.model([lat], t_tokens, [context])[0].float()
        uncond = self.model([lat], t_tokens, [context_null])[0].float()
        return uncond + self.guide_scale * (cond - uncond)

    @torch.no_grad()
    def __call__(self, latent, t_tokens, sigma: float, sigma_next: float, context, context_null, cfg_group=None, first_frame=None):
        lat = latent.to(device=self.model.device, dtype=BF16).contiguous()
        peer = False
        if cfg_group is None:
            cond = self.model([lat], t_tokens, [context])[0]
            uncond = self.model([lat], t_tokens, [context_null])[0]
        else:
            mine = self.model([lat], t_tokens, [context if cfg_group.branch == 1 else context_null])[0]
            if hasattr(cfg_group, "peer_views"):               # partner's prediction read over NVLink peer memory by the kernel below
                uncond, cond = cfg_group.peer_views(mine.contiguous())
                peer = True
            else:
                uncond, cond = cfg_group.exchange(mine)
        # x_next = x + (sigma_next - sigma) * v   <=>   x0 := v (sqrt_alpha_t = 0, sqrt_beta_t = -1), prev = 1 * x + dsigma * x0
        nxt = dense.cfg_scheduler_step(cond.contiguous(), uncond.contiguous(), lat, mode=dense.SCHED_DDIM, guidance=self.guide_scale,
                                       sqrt_alpha_t=0.0, sqrt_beta_t=-1.0, c_sample=1.0, c_x0=float(sigma_next - sigma))
        if peer:
            cfg_group.release()
        if first_frame is not None:                                     # TI2V: the first latent frame stays the encoded image
            nxt[:, :1].copy_(first_frame.to(nxt.dtype))
        return nxt
