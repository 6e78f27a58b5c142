"""Dev timing of the LayerNorm + modulation kernel at the CFG-pair shape (35 552 rows x 3072): CUDA events, bf16 in / out."""
import sys
import torch
sys.path.insert(0, ".")
from videogpa_b200 import dense
BF = torch.bfloat16
M, S, St, D = 35552, 17776, 226, 3072
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, D, device="cuda", generator=g).to(BF)
n = torch.empty_like(x)
mod = torch.randn(2, 6 * D, device="cuda", generator=g).to(BF)
ones, zeros = torch.ones(D, device="cuda", dtype=BF), torch.zeros(D, device="cuda", dtype=BF)
f = lambda: dense.layernorm_modulate(x, ones, zeros, eps=1e-5, out=n, rows_per_sample=S, text_rows=St, shift_vid=mod[:, 0:D], scale_vid=mod[:, D:2 * D],
                                     shift_txt=mod[:, 3 * D:4 * D], scale_txt=mod[:, 4 * D:5 * D], mod_stride_b=6 * D)
for _ in range(3):
    f()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"ln_modulate M={M} D={D}: {ms * 1000:.1f} us, {2 * M * D * 2 / ms / 1e6:.0f} GB/s (read + write)")
