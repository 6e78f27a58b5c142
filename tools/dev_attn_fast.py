"""Dev check + timing of the bounded-softmax head_dim-64 forward (attention_d64b_sm100.cu) on the GPU box.
usage: python tools/dev_attn_fast.py [check|bench|sweep]
  sweep re-runs `bench` in sub-processes over VGPA_ATTN_NPOLY8 (the knob is read once per process)."""
import os, subprocess, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from videogpa_b200 import dense

torch.manual_seed(0)
dev, BF = "cuda", torch.bfloat16


def errs(a, b):
    a, b = a.float(), b.float()
    mr = ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
    l2 = ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
    cs = F.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()
    return mr, l2, cs


def check(B, H, S, Skv=None, scale=1.0, lse=False):
    Skv = Skv or S
    D = H * 64
    q = (torch.randn(B, S, D, device=dev) * scale).to(BF)
    k = (torch.randn(B, Skv, D, device=dev) * scale).to(BF)
    v = torch.randn(B, Skv, D, device=dev).to(BF)
    L = torch.empty(B, H, S, device=dev, dtype=torch.float32) if lse else None
    out = dense.attention(q, k, v, H, lse=L)
    ex = dense.attention(q, k, v, H, exact=True)
    torch.cuda.synchronize()
    sp = lambda t, n: t.reshape(B, n, H, 64).transpose(1, 2).float()
    ref = F.scaled_dot_product_attention(sp(q, S), sp(k, Skv), sp(v, Skv)).transpose(1, 2).reshape(B, S, D)
    mr, l2, cs = errs(out, ref)
    mre, l2e, _ = errs(ex, ref)
    msg = f"B={B} H={H} S={S} Skv={Skv} scale={scale}: fast max-rel {mr:.3e} rel-L2 {l2:.3e} cos {cs:.6f} | exact max-rel {mre:.3e} rel-L2 {l2e:.3e}"
    if lse:
        s = torch.einsum("bhqd,bhkd->bhqk", sp(q, S), sp(k, Skv)) * 0.125 * 1.4426950408889634
        want = torch.logsumexp(s * 0.6931471805599453, dim=-1) / 0.6931471805599453
        msg += f" | lse abs err {(L - want).abs().max().item():.3e}"
    print(msg, "finite", torch.isfinite(out.float()).all().item(), flush=True)


def bench(B=2, H=48, S=17776, iters=10):
    D = H * 64
    g = torch.Generator(device=dev).manual_seed(1)
    qkv = torch.randn(B, S, 3 * D, device=dev, generator=g).to(BF)
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    out = torch.empty(B, S, D, device=dev, dtype=BF)
    fl = 4.0 * B * H * S * S * 64

    def run(tag, fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"{tag}: {ms:.3f} ms {fl / ms / 1e9:.1f} TF/s", flush=True)

    np8 = os.environ.get("VGPA_ATTN_NPOLY8", "default")
    run(f"fast NPOLY8={np8}", lambda: dense.attention(q, k, v, H, out=out))
    if os.environ.get("DEV_ATTN_ALL"):
        run("exact (online softmax)", lambda: dense.attention(q, k, v, H, out=out, exact=True))
        qh, kh, vh = (t.reshape(B, S, H, 64).transpose(1, 2) for t in (q, k, v))
        qc, kc, vc = qh.contiguous(), kh.contiguous(), vh.contiguous()
        run("torch SDPA (default backend, contiguous BHSD)", lambda: F.scaled_dot_product_attention(qc, kc, vc))
        run("torch SDPA (strided views of the fused qkv)", lambda: F.scaled_dot_product_attention(qh, kh, vh))
        from torch.nn.attention import SDPBackend, sdpa_kernel
        for be in (SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION):
            try:
                with sdpa_kernel(be):
                    run(f"torch SDPA {be.name}", lambda: F.scaled_dot_product_attention(qc, kc, vc))
            except Exception as e:  # noqa: BLE001
                print(f"torch SDPA {be.name}: unavailable ({type(e).__name__})", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "check":
        for (B, H, S) in [(1, 1, 96), (1, 1, 128), (1, 2, 256), (1, 2, 300), (2, 3, 1000), (1, 2, 4096)]:
            check(B, H, S, lse=True)
        check(1, 2, 300, 517)
        check(1, 2, 257, 49)
        check(1, 1, 1, 1)
        check(1, 2, 1000, 1000, scale=3.0)      # bound > 90: every head falls back to the exact kernel
        check(1, 1, 17776)
        check(2, 48, 2048)
    elif mode == "bench":
        bench()
    else:
        combos = [c.split(":") for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0:0", "3:0", "4:0", "4:1"])]
        for i, (np8, dbg) in enumerate(combos):
            env = dict(os.environ, VGPA_ATTN_NPOLY8=np8, VGPA_ATTN_DBG=dbg)
            if i == 0 and os.environ.get("DEV_ATTN_REF"):
                env["DEV_ATTN_ALL"] = "1"
            print(f"--- NPOLY8={np8} DBG={dbg}", flush=True)
            subprocess.run([sys.executable, __file__, "bench"], env=env, check=False)
