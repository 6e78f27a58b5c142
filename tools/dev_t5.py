"""Dev timing of the full-size T5-v1.1-XXL encoder (24 layers, 226 tokens) on the GPU box (not part of the product)."""
import sys
import torch
sys.path.insert(0, ".")
from videogpa_b200.t5 import T5Config, T5EncoderModel
enc = T5EncoderModel.random_init(T5Config(), seed=3, device="cuda")
ids = torch.randint(0, 32128, (1, 226), device="cuda")
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    out = enc(ids)[0]
    e1.record(); torch.cuda.synchronize()
    # weight bytes read once per call: 24 x (4 x 4096^2 + 3 x 4096 x 10240) x 2 B = 9.26 GB
    wb = 24 * (4 * 4096 * 4096 + 3 * 4096 * 10240) * 2
    print(f"T5-XXL encode 226 tokens: {e0.elapsed_time(e1):.2f} ms -> {wb/e0.elapsed_time(e1)/1e6:.0f} GB/s of weight traffic; "
          f"finite={torch.isfinite(out.float()).all().item()} mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
