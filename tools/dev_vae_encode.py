"""Dev timing of the full-size VAE encode (49 frames 480x720, tiled) on the GPU box (not part of the product)."""
import sys, time
import torch
sys.path.insert(0, ".")
from videogpa_b200.vae import AutoencoderKLCogVideoXEncoder, VAEDecoderConfig
enc = AutoencoderKLCogVideoXEncoder.random_init(VAEDecoderConfig(), seed=6, device="cuda")
enc.enable_tiling(); enc.enable_slicing()
x = (torch.rand(1, 3, 49, 480, 720, device="cuda") * 2 - 1).to(torch.bfloat16)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    m = enc.encode(x).latent_dist.parameters
    e1.record(); torch.cuda.synchronize()
    print(f"encode 49f 480x720 tiled: {e0.elapsed_time(e1):.1f} ms; moments {tuple(m.shape)} finite={torch.isfinite(m.float()).all().item()} "
          f"mean|m|={m.float().abs().mean().item():.3f} mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
img = (torch.rand(1, 3, 1, 480, 720, device="cuda") * 2 - 1).to(torch.bfloat16)
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    m = enc.encode(img).latent_dist.parameters
    e1.record(); torch.cuda.synchronize()
    print(f"encode 1 frame 480x720 tiled: {e0.elapsed_time(e1):.1f} ms; moments {tuple(m.shape)}", flush=True)
