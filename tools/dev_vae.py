"""Dev timing of the full-size VAE decode (49 frames 480x720) on the GPU box (not part of the product)."""
import sys, time
import torch
sys.path.insert(0, ".")
from oracle import vae_torch as V
from videogpa_b200.vae import AutoencoderKLCogVideoXDecoder, VAEDecoderConfig

cfg = V.VAEConfig()
sd = V.random_state_dict(cfg, seed=5, dtype=torch.bfloat16)
dec = AutoencoderKLCogVideoXDecoder(sd, VAEDecoderConfig(), device="cuda")
dec.enable_tiling(); dec.enable_slicing()
z = torch.randn(1, 16, 13, 60, 90, device="cuda").to(torch.bfloat16)
import os
dec.tile_streams = int(os.environ.get("VAE_STREAMS", "4"))
dec.enable_cuda_graph(os.environ.get("VAE_GRAPH", "0") == "1")
ref = None
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = dec.decode(z).sample
    enq = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    # tiled FLOPs: 9 tiles (rows 30,30,10 x cols 45,45,18 latents)
    fl = 0.0
    for hh in (30, 30, 10):
        for ww in (45, 45, 18):
            for (s, e) in dec.frame_batches(13, 2):
                fl += dec.conv_flops(e - s, hh, ww)
    if ref is None: ref = out.clone()
    print("same as first:", torch.equal(ref, out))
    print(f"decode 49f 480x720 tiled: {e0.elapsed_time(e1):.1f} ms device, {wall*1000:.1f} ms wall, {enq*1000:.1f} ms host enqueue, {fl/1e12:.1f} TFLOP conv -> {fl/e0.elapsed_time(e1)/1e9:.1f} TF/s; out {tuple(out.shape)} finite={torch.isfinite(out.float()).all().item()} mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
