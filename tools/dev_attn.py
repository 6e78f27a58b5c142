"""Dev check + timing of the attention kernel variants on the GPU box (not part of the product).
usage: python tools/dev_attn.py [check|bench]   (VGPA_ATTN_NPOLY selects the exp2 polynomial share)"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from videogpa_b200 import dense

torch.manual_seed(0)
dev, BF = "cuda", torch.bfloat16


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


def check(B, H, S, Skv=None, scale=1.0, hd=64):
    Skv = Skv or S
    D = H * hd
    q = (torch.randn(B, S, D, device=dev) * scale).to(BF)
    k = (torch.randn(B, Skv, D, device=dev) * scale).to(BF)
    v = torch.randn(B, Skv, D, device=dev).to(BF)
    out = dense.attention(q, k, v, H, head_dim=hd)
    torch.cuda.synchronize()
    sp = lambda t, n: t.reshape(B, n, H, hd).transpose(1, 2).float()
    ref = F.scaled_dot_product_attention(sp(q, S), sp(k, Skv), sp(v, Skv)).transpose(1, 2).reshape(B, S, D)
    print(f"attn hd={hd} B={B} H={H} S={S} Skv={Skv} scale={scale}: rel err {rel(out, ref):.3e} finite={torch.isfinite(out.float()).all().item()}", flush=True)


def bench(B=2, H=48, S=17776, iters=5, hd=64):
    D = H * hd
    qkv = torch.randn(B, S, 3 * D, device=dev).to(BF)
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    out = torch.empty(B, S, D, device=dev, dtype=BF)
    for _ in range(2):
        dense.attention(q, k, v, H, out=out, head_dim=hd)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dense.attention(q, k, v, H, out=out, head_dim=hd)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * B * H * S * S * hd
    print(f"hd={hd} NPOLY={os.environ.get('VGPA_ATTN_NPOLY', 'default')} attn B={B} H={H} S={S}: {ms:.3f} ms {fl / ms / 1e9:.1f} TF/s", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "check":
        for (B, H, S) in [(1, 1, 128), (1, 2, 256), (1, 2, 300), (2, 3, 1000), (1, 2, 4096)]:
            check(B, H, S)
        check(1, 2, 300, 517)
        check(1, 2, 1000, 1000, scale=3.0)      # peaky softmax: exercises the rescale path
        check(1, 1, 17776)
        for (B, H, S) in [(1, 1, 128), (1, 2, 300), (2, 3, 1000), (1, 2, 4096)]:
            check(B, H, S, hd=128)
        check(1, 2, 300, 517, hd=128)
        check(1, 2, 1000, 512, hd=128)
        check(1, 2, 1000, 1000, scale=3.0, hd=128)
        check(1, 1, 18480, hd=128)
    else:
        bench()
        bench(B=1, H=24, S=18480, hd=128)
