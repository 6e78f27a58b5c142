"""Dev: exactly one full-size tiled VAE decode (49 frames 480x720) for an ncu launch list (not part of the product)."""
import os, sys
import torch
sys.path.insert(0, ".")
from videogpa_b200.vae import AutoencoderKLCogVideoXDecoder, VAEDecoderConfig
dec = AutoencoderKLCogVideoXDecoder.random_init(VAEDecoderConfig(), seed=5, device="cuda")
dec.enable_tiling(); dec.enable_slicing()
dec.tile_streams = int(os.environ.get("VAE_STREAMS", "1"))
z = torch.randn(1, 16, 13, 60, 90, device="cuda").to(torch.bfloat16)
out = dec.decode(z).sample
torch.cuda.synchronize()
print(tuple(out.shape))
