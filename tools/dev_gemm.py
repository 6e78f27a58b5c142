"""Dev check of the tcgen05 GEMM against torch.matmul on the GPU box (not part of the product)."""
import sys, time
import torch
sys.path.insert(0, ".")
from videogpa_b200 import dense

torch.manual_seed(0)
dev = "cuda"

def check(M, N, K, epi=0):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = (torch.randn(N, device=dev) * 0.1).bfloat16()
    out = dense.linear(a, w, b, epilogue=epi)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + b.float()
    if epi == 1:
        ref = torch.nn.functional.gelu(ref.bfloat16().float(), approximate="tanh")
    err = (out.float() - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    print(f"M={M} N={N} K={K} epi={epi}: max abs err {err:.4e} rel {rel:.3e} finite={torch.isfinite(out.float()).all().item()}", flush=True)
    return rel

for (M, N, K) in [(128, 256, 64), (128, 256, 128), (256, 512, 512), (300, 64, 192), (1000, 768, 3072), (17776, 3072, 3072)]:
    check(M, N, K)
check(1000, 768, 512, epi=1)

def bench(M, N, K, epi=0, iters=10):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = (torch.randn(N, device=dev) * 0.1).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        dense.linear(a, w, b, out=out, epilogue=epi)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dense.linear(a, w, b, out=out, epilogue=epi)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    for _ in range(3):
        torch.nn.functional.linear(a, w, b)
    e0.record()
    for _ in range(iters):
        torch.nn.functional.linear(a, w, b)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"bench M={M} N={N} K={K} epi={epi}: {ms:.3f} ms {tf:.1f} TF/s | cublas {ms2:.3f} ms {2.0*M*N*K/ms2/1e9:.1f} TF/s", flush=True)

bench(35552, 3072, 3072)
bench(35552, 9216, 3072)
bench(35552, 12288, 3072, epi=1)
bench(35552, 3072, 12288)
