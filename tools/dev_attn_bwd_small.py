"""One attention forward + backward at S = 17776 with few heads, for ncu captures (dev tool)."""
import sys, torch
sys.path.insert(0, ".")
from videogpa_b200 import dense
B, H, S = 1, 8, 17776
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, S, 3 * H * 64, device="cuda", generator=g).to(torch.bfloat16)
d_out = torch.randn(B, S, H * 64, device="cuda", generator=g).to(torch.bfloat16)
q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
lse = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
out = dense.attention(q, k, v, H, lse=lse)
dq, dk, dv = dense.attention_backward(q, k, v, out, d_out, lse, H)
torch.cuda.synchronize()
print("done", torch.isfinite(dq.float()).all().item())
