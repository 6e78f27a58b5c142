"""Per-shape timing of the VAE decoder's conv launches (dev tool): CUDA events around every vgpa_conv3d call of one
tiled decode on one stream; prints count, FLOPs, time and TFLOP/s per distinct (Cin, Cout, KT, T, H, W)."""
import sys, collections
import torch
sys.path.insert(0, ".")
from oracle import vae_torch as V
from videogpa_b200.vae import AutoencoderKLCogVideoXDecoder, VAEDecoderConfig

cfg = V.VAEConfig()
sd = V.random_state_dict(cfg, seed=5, dtype=torch.bfloat16)
dec = AutoencoderKLCogVideoXDecoder(sd, VAEDecoderConfig(), device="cuda")
dec.enable_tiling(); dec.enable_slicing(); dec.tile_streams = 1
z = torch.randn(1, 16, 13, 60, 90, device="cuda").to(torch.bfloat16)
dec.decode(z)                                       # warm
recs = []
orig = dec._conv_call
def timed(cv, xpad, T, out=None, residual=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(cv, xpad, T, out=out, residual=residual); e1.record()
    recs.append(((cv.cin, cv.cout, cv.kt, T, xpad.shape[1], xpad.shape[2]), e0, e1))
    return r
dec._conv_call = timed
dec.decode(z); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for k, e0, e1 in recs:
    agg[k][0] += 1; agg[k][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"{len(recs)} conv launches, {tot:.1f} ms")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    cin, cout, kt, T, H, W = k
    fl = 2.0 * kt * 9 * cin * cout * T * H * W
    tiles = -(-H // 8) * -(-W // 16) * T * max(1, -(-cout // 256))
    print(f"  {ms:7.1f} ms {100*ms/tot:5.1f}%  n={n:4d}  {1e3*ms/n:7.1f} us  {fl*n/ms/1e9:7.0f} TF/s  Cin {cin:3d} Cout {cout:3d} KT {kt} T {T} {H}x{W}  work items {tiles}")
