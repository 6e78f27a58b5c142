import sys, os, torch
sys.path.insert(0, ".")
from videogpa_b200 import dense
BF = torch.bfloat16
M = 35552
def t(fn, it=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
res = []
for (N, K) in [(3072, 3072), (3072, 12288), (9216, 3072), (12288, 3072)]:
    a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") * 0.02).to(BF); b = torch.zeros(N, device="cuda", dtype=BF)
    out = torch.empty(M, N, device="cuda", dtype=BF)
    ms = t(lambda: dense.linear(a, w, b, out=out, epilogue=dense.EPI_BIAS))
    res.append(f"{2.0*M*N*K/ms/1e9:.0f}")
print("CLUSTER=" + os.environ.get("VGPA_GEMM_CLUSTER", "1") + " TF/s: " + " ".join(res), flush=True)
