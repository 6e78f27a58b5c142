"""Dev timing of the attention backward at the CogVideoX-5B shape (dev tool)."""
import sys, torch
sys.path.insert(0, ".")
from videogpa_b200 import dense
B, H, S = 2, 48, 17776
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, S, 3 * H * 64, device="cuda", generator=g).to(torch.bfloat16)
d_out = torch.randn(B, S, H * 64, device="cuda", generator=g).to(torch.bfloat16)
q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
lse = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
out = dense.attention(q, k, v, H, lse=lse)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    dq, dk, dv = dense.attention_backward(q, k, v, out, d_out, lse, H)
    e1.record(); torch.cuda.synchronize()
    fl = 7 * 2.0 * B * H * S * S * 64           # 3 GEMMs in the dQ kernel + 4 in the dK/dV kernel
    print(f"attention backward B={B} H={H} S={S}: {e0.elapsed_time(e1):.1f} ms, {fl/e0.elapsed_time(e1)/1e9:.0f} TFLOP/s (7 GEMM units); finite={torch.isfinite(dq.float()).all().item() and torch.isfinite(dk.float()).all().item()}", flush=True)
