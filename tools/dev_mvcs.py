"""Dev timing of vgpa_mvcs_batch at the DA3 production size (10 x 504 x 504, 128 clips per launch)."""
import math, sys, torch
sys.path.insert(0, ".")
from videogpa_b200.metrics import mvcs_batch
dev = "cuda"
N, T, H, W = 128, 10, 504, 504
gd = torch.Generator(device=dev).manual_seed(0)
depth = 2.0 + 0.5 * torch.rand(N, T, H, W, generator=gd, device=dev)
K = torch.tensor([[0.8 * W, 0, W / 2], [0, 0.8 * W, H / 2], [0, 0, 1]], device=dev).expand(N, T, 3, 3).contiguous()
E = torch.zeros(N, T, 3, 4, device=dev)
for i in range(T):
    a = math.radians(0.5 * i)
    E[:, i] = torch.tensor([[math.cos(a), 0, math.sin(a), 0.02 * i], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0]], device=dev)
for _ in range(3):
    sc = mvcs_batch(depth, K, E)
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record()
for _ in range(10):
    sc = mvcs_batch(depth, K, E)
a1.record(); torch.cuda.synchronize()
ms = a0.elapsed_time(a1) / 10
print(f"mvcs_batch {N} clips: {ms:.3f} ms/launch  {N / ms * 1e3:.0f} scores/s  {N * (T - 1) * H * W * 8 / ms / 1e6:.0f} GB/s algorithmic  score0 {sc[0].item():.12f}")
