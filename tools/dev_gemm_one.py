import sys, torch
sys.path.insert(0, ".")
from videogpa_b200 import dense
BF = torch.bfloat16
M, N, K = 35552, 12288, 3072
a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") * 0.02).to(BF); b = torch.zeros(N, device="cuda", dtype=BF)
out = torch.empty(M, N, device="cuda", dtype=BF)
for _ in range(3):
    dense.linear(a, w, b, out=out, epilogue=dense.EPI_BIAS)
torch.cuda.synchronize()
