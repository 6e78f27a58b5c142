"""Differentiable CogVideoX DiT forward for the DPO training step — SURVEY.md §8 row f-2.

The reference trains rank-64 LoRA factors on `to_q / to_k / to_v / to_out.0` of every block (PEFT `get_peft_model`,
`train/CogVideoX-5B/03_train.py:100-108`, lora_alpha 128, dropout 0) with gradient checkpointing, and calls the transformer
WITHOUT rotary embeddings (:134-139). This module is that forward on the sm_100a kernels with hand-written backward kernels:
torch.autograd only records the graph and routes gradients between `torch.autograd.Function`s whose forward AND backward are
C-ABI kernels (include/videogpa_b200.h):

    LayerNorm + adaLN modulation      vgpa_layernorm_modulate_bf16 / vgpa_layernorm_modulate_bwd_bf16
    base Linear + LoRA branch         vgpa_linear_bf16 (dgrad on pre-transposed frozen weights; LoRA dgrad / wgrad as GEMMs)
    per-head LayerNorm(64) on q, k    vgpa_head_layernorm_bf16 (forward / backward)
    attention                         vgpa_attention_bf16 (+ logsumexp) / vgpa_attention_bwd_bf16
    gated residual                    vgpa_scale_cols_bf16
    GELU(tanh)                        vgpa_gelu_tanh_bf16 (forward / backward)

Base weights are frozen and shared with the inference transformer (the DPO reference model is the same object without the
LoRA branch, so no second 11 GB copy exists); the conditioning path (time embedding, adaLN projections) gets no gradient
because it does not depend on the LoRA factors. PEFT semantics kept: y = base(x) + (lora_alpha / r) * lora_B(lora_A(x)),
A ~ kaiming_uniform(a = sqrt(5)), B = 0, fp32 master factors cast to bf16 for the GEMMs.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch.utils.checkpoint import checkpoint

from . import _lib, dense
from ._lib import LayerNormArgs

BF16 = torch.bfloat16
TARGETS = ("to_q", "to_k", "to_v", "to_out.0")


def _t(x: torch.Tensor) -> torch.Tensor:
    return x.t().contiguous()


def _t_rows(x: torch.Tensor) -> torch.Tensor:
    """[M, C] -> [C, M'] with the token dimension zero-padded to a multiple of 8: it becomes the K dimension of a weight-
    gradient GEMM (K % 8 == 0 for the 16-byte TMA row pitch); zero columns on both operands leave the product unchanged."""
    M, Cc = x.shape
    Mp = (M + 7) // 8 * 8
    if x.stride(1) != 1 or x.stride(0) % 2 or x.data_ptr() % 4:
        x = x.contiguous()
    out = torch.empty((Cc, Mp), dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().vgpa_transpose_bf16(x.data_ptr(), out.data_ptr(), M, Cc, x.stride(0), Mp, _lib.current_stream()), "vgpa_transpose_bf16")
    return out


def _ln_args(x, w, eps, seg, scale_txt, scale_vid, mod_stride_b):
    a = LayerNormArgs()
    a.x, a.rows, a.D, a.ldx = x.data_ptr(), x.shape[0], x.shape[1], x.stride(0)
    a.ln_weight = _lib.ptr(w)
    a.eps = eps
    a.rows_per_sample, a.text_rows = seg
    a.scale_txt, a.scale_vid, a.mod_stride_b = _lib.ptr(scale_txt), _lib.ptr(scale_vid), mod_stride_b
    return a


class _LNMod(torch.autograd.Function):
    """n = LN(x) * (1 + scale[b, seg]) + shift[b, seg]; gradient w.r.t. x only."""

    @staticmethod
    def forward(ctx, x, w, b, eps, seg, shift_txt, scale_txt, shift_vid, scale_vid, stride_b):
        out = dense.layernorm_modulate(x, w, b, eps=eps, rows_per_sample=seg[0], text_rows=seg[1], shift_txt=shift_txt,
                                       scale_txt=scale_txt, shift_vid=shift_vid, scale_vid=scale_vid, mod_stride_b=stride_b)
        ctx.save_for_backward(x)
        ctx.misc = (w, eps, seg, scale_txt, scale_vid, stride_b)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        w, eps, seg, scale_txt, scale_vid, stride_b = ctx.misc
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        a = _ln_args(x, w, eps, seg, scale_txt, scale_vid, stride_b)
        _lib.check(_lib.load().vgpa_layernorm_modulate_bwd_bf16(C.byref(a), dy.data_ptr(), dy.stride(0), None, 0, dx.data_ptr(),
                                                                dx.stride(0), _lib.current_stream()), "vgpa_layernorm_modulate_bwd_bf16")
        return (dx,) + (None,) * 9


class _LoRALinear(torch.autograd.Function):
    """y = a W^T + bias, plus for every adapter i: y[:, c0_i : c0_i + n_i] += s * (a A_i^T) B_i^T.
    Gradients: a, A_i, B_i (W, bias frozen; W^T is passed for the dgrad GEMM)."""

    @staticmethod
    def forward(ctx, a, w, bias, wt, s, cols, *ab):
        y = dense.linear(a, w, bias)
        us = []
        for i, (c0, n) in enumerate(cols):
            A, Bm = ab[2 * i], ab[2 * i + 1]
            u = dense.linear(a, A)                                            # [M, r]
            dense.linear(u, (Bm.float() * s).to(BF16), out=y[:, c0:c0 + n], epilogue=dense.EPI_GATE_RES)
            us.append(u)
        ctx.save_for_backward(a, *ab, *us)
        ctx.misc = (wt, s, cols)
        return y

    @staticmethod
    def backward(ctx, dy):
        wt, s, cols = ctx.misc
        n_ad = len(cols)
        saved = ctx.saved_tensors
        a, ab, us = saved[0], saved[1:1 + 2 * n_ad], saved[1 + 2 * n_ad:]
        dy = dy.contiguous()
        need_da = ctx.needs_input_grad[0]                                     # false for block 0 (its input has no graph)
        da = dense.linear(dy, wt) if need_da else None                        # dgrad through the frozen weight
        grads = []
        a_t = _t_rows(a) if n_ad else None                                    # [K, M], shared by the adapters of this input
        for i, (c0, n) in enumerate(cols):
            A, Bm, u = ab[2 * i], ab[2 * i + 1], us[i]
            dyi = dy[:, c0:c0 + n]
            du = dense.linear(dyi, _t((Bm.float() * s).to(BF16)))             # [M, r] = dy_i (s B)
            dB = dense.linear(_t_rows(dyi), _t_rows(u))                       # [n, r] = dy_i^T u
            dB = (dB.float() * s).to(BF16)
            dA = dense.linear(_t_rows(du), a_t)                               # [r, K] = du^T a
            if need_da:
                dense.linear(du, _t(A), out=da, epilogue=dense.EPI_GATE_RES)  # da += du A
            grads += [dA, dB]
        return (da, None, None, None, None, None) + tuple(grads)


class _HeadLN(torch.autograd.Function):
    """norm_q / norm_k: LayerNorm(64) per head on the q and k thirds of the fused projection; v passes through."""

    @staticmethod
    def forward(ctx, qkv, heads, lnq, lnk, eps):
        lib = _lib.load()
        M, D3 = qkv.shape
        D = D3 // 3
        out = torch.empty_like(qkv)
        for part, (w, b) in enumerate((lnq, lnk)):
            x = qkv[:, part * D:(part + 1) * D]
            o = out[:, part * D:(part + 1) * D]
            _lib.check(lib.vgpa_head_layernorm_bf16(x.data_ptr(), None, o.data_ptr(), M, heads, qkv.stride(0), 0, out.stride(0),
                                                    w.data_ptr(), b.data_ptr(), eps, 0, _lib.current_stream()), "vgpa_head_layernorm_bf16")
        out[:, 2 * D:].copy_(qkv[:, 2 * D:])
        ctx.save_for_backward(qkv)
        ctx.misc = (heads, lnq, lnk, eps)
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (qkv,) = ctx.saved_tensors
        heads, lnq, lnk, eps = ctx.misc
        dy = dy.contiguous()
        M, D3 = qkv.shape
        D = D3 // 3
        dx = torch.empty_like(qkv)
        for part, (w, b) in enumerate((lnq, lnk)):
            x = qkv[:, part * D:(part + 1) * D]
            g = dy[:, part * D:(part + 1) * D]
            o = dx[:, part * D:(part + 1) * D]
            _lib.check(lib.vgpa_head_layernorm_bf16(x.data_ptr(), g.data_ptr(), o.data_ptr(), M, heads, qkv.stride(0), dy.stride(0),
                                                    dx.stride(0), w.data_ptr(), b.data_ptr(), eps, 1, _lib.current_stream()),
                       "vgpa_head_layernorm_bf16")
        dx[:, 2 * D:].copy_(dy[:, 2 * D:])
        return dx, None, None, None, None


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, heads):
        B, S, D3 = qkv.shape
        D = D3 // 3
        lse = torch.empty((B, heads, S), dtype=torch.float32, device=qkv.device)
        out = dense.attention(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], heads, lse=lse)
        ctx.save_for_backward(qkv, out, lse)
        ctx.heads = heads
        return out

    @staticmethod
    def backward(ctx, d_out):
        qkv, out, lse = ctx.saved_tensors
        D = qkv.shape[-1] // 3
        d_qkv = torch.empty_like(qkv)
        dense.attention_backward(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], out, d_out.contiguous(), lse, ctx.heads,
                                 grads=(d_qkv[..., :D], d_qkv[..., D:2 * D], d_qkv[..., 2 * D:]))
        return d_qkv, None


def _scale_cols(x, add, seg, g_txt, g_vid, stride_b):
    out = torch.empty_like(x)
    _lib.check(_lib.load().vgpa_scale_cols_bf16(x.data_ptr(), _lib.ptr(add), out.data_ptr(), x.shape[0], x.shape[1], x.stride(0),
                                                add.stride(0) if add is not None else 0, out.stride(0), seg[0], seg[1],
                                                g_txt.data_ptr(), g_vid.data_ptr(), stride_b, _lib.current_stream()), "vgpa_scale_cols_bf16")
    return out


class _GateRes(torch.autograd.Function):
    """hidden + gate[b, seg] * branch."""

    @staticmethod
    def forward(ctx, hidden, branch, seg, g_txt, g_vid, stride_b):
        ctx.misc = (seg, g_txt, g_vid, stride_b)
        return _scale_cols(branch, hidden, seg, g_txt, g_vid, stride_b)

    @staticmethod
    def backward(ctx, d):
        seg, g_txt, g_vid, stride_b = ctx.misc
        d = d.contiguous()
        return d, _scale_cols(d, None, seg, g_txt, g_vid, stride_b), None, None, None, None


class _Gelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        out = torch.empty_like(x)
        _lib.check(_lib.load().vgpa_gelu_tanh_bf16(x.data_ptr(), None, out.data_ptr(), x.numel(), 0, _lib.current_stream()), "vgpa_gelu_tanh_bf16")
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        _lib.check(_lib.load().vgpa_gelu_tanh_bf16(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), 1, _lib.current_stream()), "vgpa_gelu_tanh_bf16")
        return dx


class LoRATrainableTransformer:
    """`get_peft_model(transformer, LoraConfig(r, lora_alpha, target_modules=[to_q, to_k, to_v, to_out.0]))` for the sm_100a DiT:
    wraps a CogVideoXTransformer3D (frozen, shared), owns the fp32 LoRA factors and runs the differentiable forward."""

    def __init__(self, transformer, r: int = 64, lora_alpha: float = 128.0, seed: int = 0, gradient_checkpointing=True):
        """gradient_checkpointing: True = every block is recomputed in the backward (the reference's
        `enable_gradient_checkpointing()`, 33 GiB at the 5B shapes); "mlp" = only the MLP half is recomputed and the
        attention half keeps its activations (sized for the 180 GB of a B200: ~125 GiB, no second attention forward);
        False = nothing is recomputed."""
        if gradient_checkpointing not in (True, False, "mlp"):
            raise RuntimeError('gradient_checkpointing must be True, False or "mlp"')
        if r % 64 != 0:
            raise RuntimeError("the LoRA rank must be a multiple of 64 (it is the N / K dimension of the LoRA GEMMs)")
        self.base = transformer
        self.config = transformer.config
        self.device = transformer.device
        self.r, self.lora_alpha, self.scaling = r, lora_alpha, lora_alpha / r
        self.gradient_checkpointing = gradient_checkpointing
        D = self.config.inner_dim
        g = torch.Generator(device=self.device).manual_seed(seed)
        bound = 1.0 / math.sqrt(D)                     # kaiming_uniform_(a = sqrt(5)) on [r, D]: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        self.lora = []
        for _ in transformer.blocks:
            layer = {}
            for m in TARGETS:
                A = ((torch.rand(r, D, device=self.device, generator=g) * 2 - 1) * bound).requires_grad_(True)
                Bm = torch.zeros(D, r, device=self.device).requires_grad_(True)
                layer[m] = (A, Bm)
            self.lora.append(layer)
        self._wt = {}

    # ------------------------------------------------------------------ PEFT-like surface
    def parameters(self):
        return [p for layer in self.lora for m in TARGETS for p in layer[m]]

    def named_parameters(self):
        for i, layer in enumerate(self.lora):
            for m in TARGETS:
                yield f"base_model.model.transformer_blocks.{i}.attn1.{m}.lora_A.weight", layer[m][0]
                yield f"base_model.model.transformer_blocks.{i}.attn1.{m}.lora_B.weight", layer[m][1]

    def lora_state_dict(self) -> dict:
        """The tensors PEFT's save_pretrained writes to adapter_model.safetensors (reference 03_train.py:287)."""
        return {k: v.detach().clone() for k, v in self.named_parameters()}

    def save_pretrained(self, path: str) -> None:
        """`model.transformer.save_pretrained(out / "final_lora")` (reference 03_train.py:287): a PEFT adapter directory —
        adapter_config.json + adapter_model.safetensors with PEFT's saved key form (`...lora_A.weight`, no adapter name) —
        that `--lora_path` of the generate CLIs (lora.merge_lora) and PEFT itself can load."""
        import json
        import os
        from safetensors.torch import save_file
        os.makedirs(path, exist_ok=True)
        cfg = {"peft_type": "LORA", "task_type": None, "base_model_name_or_path": None, "r": self.r, "lora_alpha": self.lora_alpha,
               "lora_dropout": 0.0, "target_modules": list(TARGETS), "bias": "none", "fan_in_fan_out": False, "use_dora": False,
               "use_rslora": False, "inference_mode": True, "init_lora_weights": True, "modules_to_save": None}
        with open(os.path.join(path, "adapter_config.json"), "w", encoding="utf-8") as f:
            json.dump(cfg, f, indent=2)
        save_file({k: v.detach().to("cpu").contiguous() for k, v in self.named_parameters()}, os.path.join(path, "adapter_model.safetensors"))

    def merged_delta(self, layer: int, module: str) -> torch.Tensor:
        A, Bm = self.lora[layer][module]
        return (self.scaling * (Bm @ A)).detach()

    def _wT(self, i: int):
        """Pre-transposed frozen weights of block i for the dgrad GEMMs (built once, + one model copy of HBM)."""
        w = self._wt.get(i)
        if w is None:
            blk = self.base.blocks[i]
            w = self._wt[i] = (_t(blk.w_qkv), _t(blk.w_o), _t(blk.w_ff1), _t(blk.w_ff2))
        return w

    # ------------------------------------------------------------------ one block, as its attention and MLP halves
    def _attn_half(self, x, i, emb, seg, heads):
        blk = self.base.blocks[i]
        c = self.config
        D = c.inner_dim
        B = emb.shape[0]
        S = seg[0]
        wqkv_t, wo_t, _, _ = self._wT(i)
        lay = self.lora[i]
        with torch.no_grad():
            m = dense.linear_smallm(emb, blk.n1_lw, blk.n1_lb, act_in=dense.ACT_SILU)
        n1 = _LNMod.apply(x, blk.n1_w, blk.n1_b, c.norm_eps, seg, m[:, 3 * D:4 * D], m[:, 4 * D:5 * D], m[:, 0:D], m[:, D:2 * D], 6 * D)
        ab = []
        for mod in ("to_q", "to_k", "to_v"):
            ab += [lay[mod][0].to(BF16), lay[mod][1].to(BF16)]
        qkv = _LoRALinear.apply(n1, blk.w_qkv, blk.b_qkv, wqkv_t, self.scaling, ((0, D), (D, D), (2 * D, D)), *ab)
        qkv = _HeadLN.apply(qkv, heads, blk.lnq, blk.lnk, 1e-6)
        att = _Attention.apply(qkv.view(B, S, 3 * D), heads).view(B * S, D)
        o = _LoRALinear.apply(att, blk.w_o, blk.b_o, wo_t, self.scaling, ((0, D),), lay["to_out.0"][0].to(BF16), lay["to_out.0"][1].to(BF16))
        return _GateRes.apply(x, o, seg, m[:, 5 * D:6 * D], m[:, 2 * D:3 * D], 6 * D)

    def _mlp_half(self, x, i, emb, seg):
        blk = self.base.blocks[i]
        c = self.config
        D = c.inner_dim
        _, _, wff1_t, wff2_t = self._wT(i)
        with torch.no_grad():
            m2 = dense.linear_smallm(emb, blk.n2_lw, blk.n2_lb, act_in=dense.ACT_SILU)
        n2 = _LNMod.apply(x, blk.n2_w, blk.n2_b, c.norm_eps, seg, m2[:, 3 * D:4 * D], m2[:, 4 * D:5 * D], m2[:, 0:D], m2[:, D:2 * D], 6 * D)
        pre = _LoRALinear.apply(n2, blk.w_ff1, blk.b_ff1, wff1_t, self.scaling, ())
        act = _Gelu.apply(pre)
        f = _LoRALinear.apply(act, blk.w_ff2, blk.b_ff2, wff2_t, self.scaling, ())
        return _GateRes.apply(x, f, seg, m2[:, 5 * D:6 * D], m2[:, 2 * D:3 * D], 6 * D)

    def _block(self, x, i, emb, seg, heads):
        return self._mlp_half(self._attn_half(x, i, emb, seg, heads), i, emb, seg)

    # ------------------------------------------------------------------ forward
    def forward(self, hidden_states, encoder_hidden_states, timestep, num_layers: int | None = None):
        """hidden_states [B, F, C, H, W], encoder_hidden_states [B, St, 4096], timestep [B] -> sample [B, F, C, H, W] (bf16) with an
        autograd graph reaching the LoRA factors. No rotary embedding, as in the reference's training call."""
        t = self.base
        c = self.config
        pt = c.patch_size_t or 1                                   # CogVideoX1.5 (train/CogVideoX1.5-5B/03_train.py): temporal patches of 2
        B, Fr, Cc, H, W = hidden_states.shape
        if Fr % pt != 0:
            raise RuntimeError(f"the number of latent frames ({Fr}) must be a multiple of patch_size_t ({pt}); the training step trims it")
        p, D, heads = c.patch_size, c.inner_dim, c.num_attention_heads
        dev = self.device
        with torch.no_grad():
            hs = hidden_states.to(device=dev, dtype=BF16).contiguous()
            enc_in = encoder_hidden_states.to(device=dev, dtype=BF16).contiguous()
            St = enc_in.shape[1]
            hw = (H // p) * (W // p)
            Sv = (Fr // pt) * hw
            S = St + Sv
            ts = torch.as_tensor(timestep, device=dev).reshape(-1).to(torch.float32)
            if ts.numel() == 1 and B > 1:
                ts = ts.expand(B).contiguous()
            t_emb = dense.timestep_embedding(ts, D)
            e1 = dense.linear_smallm(t_emb, t.t1_w, t.t1_b)
            emb = dense.linear_smallm(e1, t.t2_w, t.t2_b, act_in=dense.ACT_SILU)
            x0 = torch.empty((B, S, D), dtype=BF16, device=dev)
            patches = dense.patchify(hs.view(B * Fr, Cc, H, W))
            if pt > 1:                                             # [B, Fr/pt, pt, hw, Cpp] -> [.., hw, pt*Cpp] (the weight columns were permuted at load)
                patches = patches.view(B, Fr // pt, pt, hw, -1).permute(0, 1, 3, 2, 4).reshape(B * Sv, -1).contiguous()
            epi = dense.EPI_BIAS
            if t.pos_embedding is not None:                        # CogVideoX-5B-I2V: learned positional embedding (frozen)
                if t.pos_embedding.shape[1] < S:
                    raise RuntimeError("pos_embedding is shorter than the token sequence")
                x0.copy_(t.pos_embedding[:, :S].expand(B, S, D))
                epi = dense.EPI_GATE_RES
            for b in range(B):
                dense.linear(enc_in[b], t.text_w, t.text_b, out=x0[b, :St], epilogue=epi)
                dense.linear(patches[b * Sv:(b + 1) * Sv], t.patch_w, t.patch_b, out=x0[b, St:], epilogue=epi)
            mo = dense.linear_smallm(emb, t.no_lw, t.no_lb, act_in=dense.ACT_SILU)
        seg = (S, St)
        x = x0.view(B * S, D)
        L = c.num_layers if num_layers is None else num_layers
        mode = self.gradient_checkpointing if torch.is_grad_enabled() else False
        for i in range(L):
            if mode == "mlp":        # keep the attention half's activations (2.2 GB per block at the 5B shapes), recompute the MLP half
                x = self._attn_half(x, i, emb, seg, heads)
                x = checkpoint(self._mlp_half, x, i, emb, seg, use_reentrant=False)
            elif mode:
                x = checkpoint(self._block, x, i, emb, seg, heads, use_reentrant=False)
            else:
                x = self._block(x, i, emb, seg, heads)
        n = _LNMod.apply(x, t.nf_w, t.nf_b, c.norm_eps, (0, 0), None, None, None, None, 0)
        y = _LNMod.apply(n, t.no_w, t.no_b, c.norm_eps, seg, mo[:, 0:D], mo[:, D:2 * D], mo[:, 0:D], mo[:, D:2 * D], 2 * D)
        if "po_t" not in self._wt:
            self._wt["po_t"] = _t(t.po_w)
        tok = _LoRALinear.apply(y, t.po_w, t.po_b, self._wt["po_t"], self.scaling, ())             # [B*S, p*p*C]
        tok = tok.view(B, S, -1)[:, St:]                                                             # video rows
        Co = c.out_channels
        # proj_out features are (pt, c, ph, pw) (rows permuted at load for pt > 1): token (f, hg, wg) -> frames f*pt .. f*pt + pt - 1
        out = tok.reshape(B, Fr // pt, H // p, W // p, pt, Co, p, p).permute(0, 1, 4, 5, 2, 6, 3, 7).reshape(B, Fr, Co, H, W)
        return out

    __call__ = forward
