"""Dev: a few launches of the MVCS and DPO-loss kernels at their bench shapes for `ncu --set full` (not part of the product)."""
import sys
import torch
sys.path.insert(0, ".")
import bench_legs
from videogpa_b200.loss import create_loss_strategy
from videogpa_b200.metrics import mvcs_batch
dev = torch.device("cuda", 0)
depth, K, E = bench_legs.scorer_inputs(dev, 32, 10, 504, 504)
for _ in range(3):
    mvcs_batch(depth, K, E)
g = torch.Generator(device=dev).manual_seed(0)
six = [torch.randn(1, 13, 16, 60, 90, device=dev, generator=g) for _ in range(6)]
loss = create_loss_strategy("dpo", beta=1.0)
for _ in range(3):
    out = loss(*six)
torch.cuda.synchronize()
print(float(out.loss))
