"""Dev timing of the full-size Wan2.2 TI2V-5B denoise step (81 frames 1280x704: latent [48,21,44,80], S = 18480)."""
import sys
import torch
sys.path.insert(0, ".")
from videogpa_b200.wan import WanConfig, WanDenoiseStep, WanTransformer3D, flow_sigmas

cfg = WanConfig.ti2v_5b()
model = WanTransformer3D.random_init(cfg, seed=21, device="cuda")
step = WanDenoiseStep(model, guide_scale=5.0)
g = torch.Generator(device="cuda").manual_seed(0)
lat = torch.randn(48, 21, 44, 80, device="cuda", generator=g).to(torch.bfloat16)
ctx = torch.randn(512, 4096, device="cuda", generator=g).to(torch.bfloat16)
ctx0 = torch.zeros(1, 4096, device="cuda", dtype=torch.bfloat16)
S, hw = 21 * 22 * 40, 22 * 40
sig = flow_sigmas(50, 5.0)
t = torch.full((1, S), sig[0] * 1000); t[:, :hw] = 0
for _ in range(2):
    x = step(lat, t, sig[0], sig[1], ctx, ctx0, first_frame=lat[:, :1])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for i in range(n):
    x = step(x, t, sig[i], sig[i + 1], ctx, ctx0, first_frame=lat[:, :1])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 2 * model.flops_per_forward(S)
print(f"Wan2.2-TI2V-5B denoise step (2 forwards + CFG + Euler), S={S}: {ms:.1f} ms/step, {S/ (ms/1000):.0f} latent tokens/s, {fl/ms/1e9:.0f} TF/s, finite={torch.isfinite(x.float()).all().item()}")
