"""PEFT LoRA adapter loading and merge for the `--lora_path` surface.

Reference: generate/CogVideoX-5B.py:24-31 (`PeftModel.from_pretrained(...).merge_and_unload()`),
scaling overrides in generate/CogVideoX1.5-5B.py:32-35 (absolute) and generate/Wan2.2-TI2V-5B.py:66-70
(multiplicative), adapter format checkpoints/*/adapter_config.json + adapter_model.safetensors
(download_ckpt.py:40-61). Merge rule: W <- W + (lora_alpha / r) * B @ A for attn1.{to_q,to_k,to_v,to_out.0}
(fp32 product, one rounding to bf16).

The product runs on the tcgen05 GEMM: the fp32 adapter matrices are split into bf16 hi + lo parts and
concatenated along K so that one GEMM (K = 3r) accumulates B_hi A_hi + B_hi A_lo + B_lo A_hi in fp32
(relative error ~2^-16 of the delta), and the VGPA_EPI_ACCUM epilogue adds it to W with one rounding.
"""
from __future__ import annotations

import json
import os
import re

import torch

from . import dense

_KEY = re.compile(r"(?:base_model\.model\.)?transformer_blocks\.(\d+)\.attn1\.(to_q|to_k|to_v|to_out\.0)\.lora_(A|B)(?:\.default)?\.weight$")
# Wan2.2 adapters target q, k, v, o of self_attn and cross_attn (train/Wan2.2-TI2V-5B/03_train.py:82)
_KEY_WAN = re.compile(r"(?:base_model\.model\.)?blocks\.(\d+)\.((?:self|cross)_attn\.[qkvo])\.lora_(A|B)(?:\.default)?\.weight$")


def read_adapter(lora_path: str):
    """-> (config dict, {(layer, module): (A [r, in], B [out, r])}) from a PEFT adapter directory."""
    cfg_file = os.path.join(lora_path, "adapter_config.json")
    if not os.path.isfile(cfg_file):
        raise RuntimeError(f"{cfg_file} not found: not a PEFT adapter directory")
    with open(cfg_file, "r", encoding="utf-8") as f:
        cfg = json.load(f)
    if cfg.get("peft_type", "LORA") != "LORA" or cfg.get("use_dora") or cfg.get("use_rslora") or cfg.get("fan_in_fan_out"):
        raise RuntimeError("only plain LoRA adapters (no DoRA / rsLoRA / fan_in_fan_out) are supported")
    st_file = os.path.join(lora_path, "adapter_model.safetensors")
    if os.path.isfile(st_file):
        from safetensors.torch import load_file
        tensors = load_file(st_file)
    elif os.path.isfile(os.path.join(lora_path, "adapter_model.bin")):
        tensors = torch.load(os.path.join(lora_path, "adapter_model.bin"), map_location="cpu")
    else:
        raise RuntimeError(f"no adapter_model.safetensors under {lora_path}")
    pairs: dict = {}
    for k, v in tensors.items():
        m = _KEY.search(k) or _KEY_WAN.search(k)
        if not m:
            continue
        key = (int(m.group(1)), m.group(2))
        pairs.setdefault(key, {})[m.group(3)] = v
    out = {}
    for key, ab in pairs.items():
        if "A" not in ab or "B" not in ab:
            raise RuntimeError(f"adapter is missing lora_A or lora_B for {key}")
        out[key] = (ab["A"], ab["B"])
    if not out:
        raise RuntimeError("no attention LoRA tensors (CogVideoX attn1.* or Wan {self,cross}_attn.*) found in the adapter")
    return cfg, out


def _split_bf16(t: torch.Tensor):
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


def _check_shapes(layer, module, A, B, W):
    if B.shape[0] != W.shape[0] or A.shape[1] != W.shape[1] or A.shape[0] != B.shape[1]:
        raise RuntimeError(f"LoRA shape mismatch at layer {layer} {module}: A {tuple(A.shape)} B {tuple(B.shape)} W {tuple(W.shape)}")


def _delta_operands(A: torch.Tensor, B: torch.Tensor):
    """fp32 A [r, in], B [out, r] -> bf16 GEMM operands ([out, 3r], [in, 3r]) whose product is B @ A to ~2^-16."""
    Bh, Bl = _split_bf16(B)
    Ah, Al = _split_bf16(A.t().contiguous())                                 # [in, r]
    a_op = torch.cat([Bh, Bh, Bl], dim=1).contiguous()                       # [out, 3r]
    w_op = torch.cat([Ah, Al, Ah], dim=1).contiguous()                       # [in, 3r]  (N = in features)
    return a_op, w_op


def merge_lora(transformer, lora_path: str, scaling: float | None = None, weight: float | None = None) -> int:
    """Merge the adapter into `transformer` in place. scaling = absolute override; weight multiplies alpha/r.

    Returns the number of merged modules.
    """
    cfg, pairs = read_adapter(lora_path)
    base = float(cfg["lora_alpha"]) / float(cfg["r"])
    s = base if scaling is None else float(scaling)
    if weight is not None:
        s = s * float(weight)
    dev = transformer.device
    n = 0
    for (layer, module), (A, B) in sorted(pairs.items()):
        if layer >= len(transformer.blocks):
            raise RuntimeError(f"adapter targets layer {layer}, the model has {len(transformer.blocks)}")
        W = transformer.attention_weight(layer, module)                      # [out, in] bf16 view, updated in place
        A = A.to(device=dev, dtype=torch.float32)                            # [r, in]
        B = B.to(device=dev, dtype=torch.float32)                            # [out, r]
        _check_shapes(layer, module, A, B, W)
        a_op, w_op = _delta_operands(A, B)
        dense.linear(a_op, w_op, None, out=W, epilogue=dense.EPI_ACCUM, alpha=s)
        n += 1
    return n


class AttachedLoRA:
    """An adapter whose strength can be changed or removed after loading.

    Reference: replicate.py:208-213 keeps the PEFT adapter unmerged and, per work item, sets
    `module.scaling[name] = lora_weight * lora_alpha / r` on every LoRA layer before generating. Here the targeted base
    weights (3.2 GB bf16 for CogVideoX-5B: a rounding error of the 180 GB) are kept pristine on the device next to the split
    adapter operands, and every change of strength rebuilds W = round_bf16(W_base + s * B @ A) from the pristine copy with
    the same accumulate-epilogue GEMM as `merge_lora`: one rounding, no drift however often the strength changes, and the
    denoise step keeps running on plain fused weights. `set_scaling(s)` with s equal to merge_lora's gives the same bits as
    `merge_lora`; `unmerge()` restores the base weights exactly.
    """

    def __init__(self, transformer, lora_path: str):
        self.transformer = transformer
        self.config, pairs = read_adapter(lora_path)
        self.r = int(self.config["r"])
        self.lora_alpha = float(self.config["lora_alpha"])
        self.scaling = 0.0                                                   # what is currently folded into the weights
        dev = transformer.device
        self._items = []
        for (layer, module), (A, B) in sorted(pairs.items()):
            if layer >= len(transformer.blocks):
                raise RuntimeError(f"adapter targets layer {layer}, the model has {len(transformer.blocks)}")
            W = transformer.attention_weight(layer, module)
            A = A.to(device=dev, dtype=torch.float32)
            B = B.to(device=dev, dtype=torch.float32)
            _check_shapes(layer, module, A, B, W)
            a_op, w_op = _delta_operands(A, B)
            self._items.append((W, W.clone(), a_op, w_op))

    def __len__(self) -> int:
        return len(self._items)

    def set_scaling(self, scaling: float) -> None:
        """Absolute strength s (PEFT's `module.scaling`): W <- W_base + s * B @ A."""
        s = float(scaling)
        for W, base, a_op, w_op in self._items:
            W.copy_(base)
            if s != 0.0:
                dense.linear(a_op, w_op, None, out=W, epilogue=dense.EPI_ACCUM, alpha=s)
        self.scaling = s

    def set_weight(self, lora_weight: float) -> None:
        """replicate.py:208-213: scaling = lora_weight * lora_alpha / r."""
        self.set_scaling(float(lora_weight) * self.lora_alpha / self.r)

    def unmerge(self) -> None:
        """Back to the base weights (bit-exact)."""
        self.set_scaling(0.0)


def attach_lora(transformer, lora_path: str, weight: float | None = 1.0) -> AttachedLoRA:
    """Load an adapter re-scalably; `weight` (default 1.0 = alpha / r, what PEFT loads with) is applied at once, None leaves
    the base weights untouched until set_scaling / set_weight is called."""
    h = AttachedLoRA(transformer, lora_path)
    if weight is not None:
        h.set_weight(weight)
    return h
