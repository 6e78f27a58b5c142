"""T5 (v1.1) encoder on the sm_100a kernels — the prompt encoder of the reference's pipelines, SURVEY.md §8 row f-4.

Drop-in for the part of transformers' `T5EncoderModel` the reference uses: `text_encoder(input_ids)[0]` with 226
max-length-padded tokens and NO attention mask (`train/CogVideoX-5B/02_encode.py:69-84`; diffusers' `_get_t5_prompt_embeds`
does the same behind `generate/CogVideoX-5B.py:72-77`). State-dict names are transformers' (`shared.weight`,
`encoder.block.{i}.layer.0.SelfAttention.{q,k,v,o}.weight`, `...relative_attention_bias.weight` on block 0,
`encoder.block.{i}.layer.1.DenseReluDense.{wi_0,wi_1,wo}.weight`, `layer_norm.weight`, `encoder.final_layer_norm.weight`).

Per block: T5LayerNorm (RMS, no bias) -> fused q|k|v GEMM -> attention with the shared relative-position bias and no
1/sqrt(d) scaling (vgpa_t5_attention_bf16) -> o GEMM with the residual add in its epilogue -> T5LayerNorm -> wi_0 GEMM with
GELU(tanh) epilogue, wi_1 GEMM, gated product (vgpa_gated_mul_bf16) -> wo GEMM + residual. The bias table [H, S, S] is built
once per sequence length from `relative_attention_bias` with transformers' bucket rule (integer, bit-exact).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace

import torch

from . import _lib, dense

BF16 = torch.bfloat16


@dataclass
class T5Config:
    """transformers.T5Config fields used by the encoder; defaults = google/t5-v1_1-xxl (CogVideoX `text_encoder/`)."""
    vocab_size: int = 32128
    d_model: int = 4096
    d_kv: int = 64
    d_ff: int = 10240
    num_layers: int = 24
    num_heads: int = 64
    relative_attention_num_buckets: int = 32
    relative_attention_max_distance: int = 128
    layer_norm_epsilon: float = 1e-6
    per_layer_relative_bias: bool = False   # umT5 (transformers UMT5EncoderModel; Wan2.2's text encoder): every block has its own table

    @classmethod
    def umt5_xxl(cls) -> "T5Config":
        """google/umt5-xxl, the text encoder of Wan2.2 (generate/Wan2.2-TI2V-5B.py via the Wan repository): T5 v1.1 geometry, 256 384-token
        vocabulary, one relative-position table per block."""
        return cls(vocab_size=256384, per_layer_relative_bias=True)


def relative_position_bucket(relative_position: torch.Tensor, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """transformers T5Attention._relative_position_bucket, bidirectional (encoder): half the buckets per sign, exact
    buckets below num_buckets/4, log-spaced up to max_distance."""
    num_buckets //= 2
    buckets = (relative_position > 0).to(torch.long) * num_buckets
    rp = torch.abs(relative_position)
    max_exact = num_buckets // 2
    is_small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rp, large)


def position_buckets(S: int, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """[S, S] bucket of (memory_position - query_position), T5Attention.compute_bias."""
    ctx = torch.arange(S, dtype=torch.long)[:, None]
    mem = torch.arange(S, dtype=torch.long)[None, :]
    return relative_position_bucket(mem - ctx, num_buckets, max_distance)


def wan_umt5_to_transformers_names(sd: dict) -> dict:
    """Rename the parameters of Wan2.2's own umT5 encoder module (`models_t5_umt5-xxl-enc-bf16.pth`, wan/modules/t5.py of the
    un-vendored Wan repository) to transformers' UMT5EncoderModel names, which this encoder loads. **[recalled]** — the Wan
    repository is not in the reference tree, so the source names below could not be checked here; a checkpoint with other names
    fails loudly (missing key) instead of loading wrongly.
        token_embedding.weight                      -> shared.weight
        blocks.{i}.norm1.weight / norm2.weight      -> encoder.block.{i}.layer.0 / layer.1 .layer_norm.weight
        blocks.{i}.attn.{q,k,v,o}.weight            -> encoder.block.{i}.layer.0.SelfAttention.{q,k,v,o}.weight
        blocks.{i}.pos_embedding.embedding.weight   -> encoder.block.{i}.layer.0.SelfAttention.relative_attention_bias.weight
        blocks.{i}.ffn.gate.0.weight / fc1 / fc2    -> encoder.block.{i}.layer.1.DenseReluDense.wi_0 / wi_1 / wo .weight
        norm.weight                                 -> encoder.final_layer_norm.weight"""
    import re
    out = {}
    rules = [(r"^token_embedding\.weight$", "shared.weight"), (r"^norm\.weight$", "encoder.final_layer_norm.weight"),
             (r"^blocks\.(\d+)\.norm1\.weight$", r"encoder.block.\1.layer.0.layer_norm.weight"),
             (r"^blocks\.(\d+)\.norm2\.weight$", r"encoder.block.\1.layer.1.layer_norm.weight"),
             (r"^blocks\.(\d+)\.attn\.([qkvo])\.weight$", r"encoder.block.\1.layer.0.SelfAttention.\2.weight"),
             (r"^blocks\.(\d+)\.pos_embedding\.embedding\.weight$", r"encoder.block.\1.layer.0.SelfAttention.relative_attention_bias.weight"),
             (r"^blocks\.(\d+)\.ffn\.gate\.0\.weight$", r"encoder.block.\1.layer.1.DenseReluDense.wi_0.weight"),
             (r"^blocks\.(\d+)\.ffn\.fc1\.weight$", r"encoder.block.\1.layer.1.DenseReluDense.wi_1.weight"),
             (r"^blocks\.(\d+)\.ffn\.fc2\.weight$", r"encoder.block.\1.layer.1.DenseReluDense.wo.weight")]
    for k, v in sd.items():
        for pat, rep in rules:
            if re.match(pat, k):
                out[re.sub(pat, rep, k)] = v
                break
        else:
            raise RuntimeError(f"unexpected parameter {k!r} in a Wan umT5 checkpoint")
    return out


class T5EncoderOutput(tuple):
    """`out[0]` / `out.last_hidden_state` like transformers' BaseModelOutput."""

    @property
    def last_hidden_state(self):
        return self[0]


class T5EncoderModel:
    def __init__(self, config: T5Config, state_dict: dict, device="cuda"):
        self.config = c = config
        self.device = torch.device(device)
        self.dtype = BF16
        if c.d_kv != 64:
            raise RuntimeError("only d_kv = 64 is supported")
        self.inner = c.num_heads * c.d_kv
        if c.d_model % 256 or c.d_model > 4096:
            raise RuntimeError("d_model must be a multiple of 256, at most 4096")
        sd = state_dict

        def get(name):
            if name not in sd:
                raise RuntimeError(f"state dict is missing {name}")
            return sd[name]

        def w16(name):
            return get(name).to(device=self.device, dtype=BF16).contiguous()

        def w32(name):                                   # T5LayerNorm weights: the kernel multiplies in fp32 after the bf16 cast
            return get(name).to(device=self.device, dtype=BF16).float().contiguous()

        emb = "shared.weight" if "shared.weight" in sd else "encoder.embed_tokens.weight"
        self.embed = w16(emb)
        self.rel_bias = w16("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight")    # [buckets, H]
        self.blocks = []
        for i in range(c.num_layers):
            a = f"encoder.block.{i}.layer.0."
            f = f"encoder.block.{i}.layer.1."
            b = SimpleNamespace()
            b.rel_bias = w16(a + "SelfAttention.relative_attention_bias.weight") if c.per_layer_relative_bias else None
            b.ln0 = w32(a + "layer_norm.weight")
            b.wqkv = torch.cat([w16(a + "SelfAttention.q.weight"), w16(a + "SelfAttention.k.weight"),
                                w16(a + "SelfAttention.v.weight")], 0).contiguous()
            b.wo = w16(a + "SelfAttention.o.weight")
            b.ln1 = w32(f + "layer_norm.weight")
            b.wi0 = w16(f + "DenseReluDense.wi_0.weight")
            b.wi1 = w16(f + "DenseReluDense.wi_1.weight")
            b.wo2 = w16(f + "DenseReluDense.wo.weight")
            self.blocks.append(b)
        self.final_ln = w32("encoder.final_layer_norm.weight")
        self._bias_cache: dict = {}
        self.use_cuda_graph = True
        self._graphs: dict = {}

    @classmethod
    def random_init(cls, config: T5Config | None = None, seed: int = 3, device="cuda") -> "T5EncoderModel":
        """Seeded synthetic weights with transformers' parameter names (no checkpoint is reachable)."""
        c = config or T5Config()
        dev = torch.device(device)
        g = torch.Generator(device=dev).manual_seed(seed)
        inner = c.num_heads * c.d_kv

        def rnd(*shape, std):
            return (torch.randn(*shape, generator=g, device=dev) * std).to(BF16)

        sd = {"shared.weight": rnd(c.vocab_size, c.d_model, std=1.0),
              "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight": rnd(c.relative_attention_num_buckets, c.num_heads, std=0.5),
              "encoder.final_layer_norm.weight": (1.0 + 0.1 * torch.randn(c.d_model, generator=g, device=dev)).to(BF16)}
        for i in range(c.num_layers):
            a, f = f"encoder.block.{i}.layer.0.", f"encoder.block.{i}.layer.1."
            if c.per_layer_relative_bias and i > 0:
                sd[a + "SelfAttention.relative_attention_bias.weight"] = rnd(c.relative_attention_num_buckets, c.num_heads, std=0.5)
            sd[a + "layer_norm.weight"] = (1.0 + 0.1 * torch.randn(c.d_model, generator=g, device=dev)).to(BF16)
            sd[a + "SelfAttention.q.weight"] = rnd(inner, c.d_model, std=(c.d_model * c.d_kv) ** -0.5)
            sd[a + "SelfAttention.k.weight"] = rnd(inner, c.d_model, std=c.d_model ** -0.5)
            sd[a + "SelfAttention.v.weight"] = rnd(inner, c.d_model, std=c.d_model ** -0.5)
            sd[a + "SelfAttention.o.weight"] = rnd(c.d_model, inner, std=inner ** -0.5)
            sd[f + "layer_norm.weight"] = (1.0 + 0.1 * torch.randn(c.d_model, generator=g, device=dev)).to(BF16)
            sd[f + "DenseReluDense.wi_0.weight"] = rnd(c.d_ff, c.d_model, std=c.d_model ** -0.5)
            sd[f + "DenseReluDense.wi_1.weight"] = rnd(c.d_ff, c.d_model, std=c.d_model ** -0.5)
            sd[f + "DenseReluDense.wo.weight"] = rnd(c.d_model, c.d_ff, std=c.d_ff ** -0.5)
        return cls(c, sd, device=device)

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ pieces
    def position_bias(self, S: int, layer: int = 0) -> torch.Tensor:
        """[H, S, S] bf16: relative_attention_bias(bucket(j - i)) permuted like T5Attention.compute_bias. T5 shares block 0's table
        over all blocks; umT5 (`per_layer_relative_bias`) gives every block its own (UMT5Attention.compute_bias per layer)."""
        key = (S, layer if self.config.per_layer_relative_bias else 0)
        b = self._bias_cache.get(key)
        if b is None:
            c = self.config
            table = self.blocks[layer].rel_bias if c.per_layer_relative_bias else self.rel_bias
            bk = position_buckets(S, c.relative_attention_num_buckets, c.relative_attention_max_distance).to(self.device)
            b = self._bias_cache[key] = table[bk].permute(2, 0, 1).contiguous()
        return b

    def _rmsnorm(self, x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
        y = x.clone()
        lib = _lib.load()
        _lib.check(lib.vgpa_rmsnorm_rope_bf16(y.data_ptr(), y.shape[0], y.shape[1], y.stride(0), w.data_ptr(),
                                              self.config.layer_norm_epsilon, None, None, 0, 0, _lib.current_stream()),
                   "vgpa_rmsnorm_rope_bf16")
        return y

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, input_ids: torch.Tensor, attention_mask: torch.Tensor | None = None, **kw) -> T5EncoderOutput:
        lib = _lib.load()
        if not isinstance(input_ids, torch.Tensor) or not input_ids.is_cuda:
            raise RuntimeError("input_ids must be a CUDA tensor (no CPU fallback exists)")
        if input_ids.dim() != 2:
            raise RuntimeError("input_ids must be [B, S]")
        if attention_mask is not None and not bool(attention_mask.to(torch.bool).all()):
            if not self.config.per_layer_relative_bias:
                raise RuntimeError("attention masks with padding are not supported (the reference passes none)")
            return self._masked(input_ids, attention_mask)
        B, S = input_ids.shape
        if S > 512:
            raise RuntimeError("sequence length above 512 is not supported")
        lo, hi = torch.aminmax(input_ids)
        if int(lo) < 0 or int(hi) >= self.embed.shape[0]:
            raise RuntimeError("token id out of range")
        ids = input_ids.to(torch.long)
        if not self.use_cuda_graph:
            return T5EncoderOutput((self._forward(ids),))
        # 24 blocks x 11 short launches on 226 rows: issued from Python the host is the limiter, so replay one graph per (B, S)
        ent = self._graphs.get((B, S))
        if ent is None:
            self._forward(ids)                                   # eager pass: bias table and lazy state exist before capture
            torch.cuda.synchronize(self.device)
            static_ids = ids.clone()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self._forward(static_ids)
            ent = self._graphs[(B, S)] = (g, static_ids, static_out)
        g, static_ids, static_out = ent
        static_ids.copy_(ids)
        g.replay()
        return T5EncoderOutput((static_out.clone(),))

    def _masked(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> T5EncoderOutput:
        """umT5 as Wan2.2 calls it: prompts padded to text_len with a mask, the output trimmed to the true lengths. Masked keys do
        not take part in attention and the position bias is relative, so the rows of the real tokens equal an unmasked run over
        the unpadded prefix: every sample is encoded at its own length; rows at padded positions are returned as zeros (transformers
        computes values there that every caller discards)."""
        m = attention_mask.to(torch.bool)
        lens = m.sum(dim=1)
        if not bool((m == (torch.arange(m.shape[1], device=m.device)[None] < lens[:, None])).all()) or int(lens.min()) == 0:
            raise RuntimeError("only right-padded attention masks with at least one token per sample are supported")
        B, S = input_ids.shape
        out = torch.zeros((B, S, self.config.d_model), dtype=BF16, device=self.device)
        for b in range(B):
            L = int(lens[b])
            out[b, :L] = self(input_ids[b:b + 1, :L])[0][0]
        return T5EncoderOutput((out,))

    def _forward(self, ids: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        B, S = ids.shape
        c = self.config
        x = self.embed.index_select(0, ids.reshape(-1)).contiguous()                               # [B*S, d_model]
        M, inner = B * S, self.inner
        stream = _lib.current_stream
        for li, b in enumerate(self.blocks):
            bias = self.position_bias(S, li)
            n = self._rmsnorm(x, b.ln0)
            qkv = dense.linear(n, b.wqkv)
            ctx = torch.empty((M, inner), dtype=BF16, device=self.device)
            esz = qkv.element_size()
            _lib.check(lib.vgpa_t5_attention_bf16(qkv.data_ptr(), qkv.data_ptr() + inner * esz, qkv.data_ptr() + 2 * inner * esz,
                                                  bias.data_ptr(), ctx.data_ptr(), B, c.num_heads, S, qkv.stride(0), ctx.stride(0),
                                                  stream()), "vgpa_t5_attention_bf16")
            dense.linear(ctx, b.wo, out=x, epilogue=dense.EPI_GATE_RES)                            # x += ctx @ Wo^T
            n = self._rmsnorm(x, b.ln1)
            g = dense.linear(n, b.wi0, epilogue=dense.EPI_BIAS_GELU)                               # gelu_new(wi_0 x)
            l = dense.linear(n, b.wi1)
            _lib.check(lib.vgpa_gated_mul_bf16(g.data_ptr(), l.data_ptr(), g.data_ptr(), M, c.d_ff, g.stride(0), l.stride(0),
                                               g.stride(0), stream()), "vgpa_gated_mul_bf16")
            dense.linear(g, b.wo2, out=x, epilogue=dense.EPI_GATE_RES)
        return self._rmsnorm(x, self.final_ln).view(B, S, c.d_model)
