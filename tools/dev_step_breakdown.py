"""Dev: CUDA-event breakdown of one full-size CogVideoX-5B denoise step by kernel family (not part of the product)."""
import sys, collections
import torch
sys.path.insert(0, ".")
from videogpa_b200 import dense
import videogpa_b200.transformer as tr
from videogpa_b200.pipeline import CogVideoXDenoisePipeline
from videogpa_b200.schedulers import CogVideoXDDIMScheduler
from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig

cfg = TransformerConfig.cogvideox_5b()
model = CogVideoXTransformer3D.random_init(cfg, seed=1234, device="cuda")
sched = CogVideoXDDIMScheduler(); ts = sched.set_timesteps(50)
pipe = CogVideoXDenoisePipeline(model, sched)
lat = pipe.prepare_latents(1, 49, 480, 720, generator=torch.Generator(device="cuda").manual_seed(42))
pe = torch.randn(2, 226, 4096, device="cuda").to(torch.bfloat16)
rope = pipe.rotary(13, 60, 90)
for i in range(3):
    lat = pipe.denoise_step(lat, pe, int(ts[i]), 6.0, rope)
events = collections.defaultdict(list)
def wrap(name, fn, key=None):
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        kk = name if key is None else name + ":" + str(key(a, k))
        events[kk].append((e0, e1)); return r
    return f
orig = {n: getattr(dense, n) for n in ("linear", "attention", "layernorm_modulate", "linear_smallm", "patchify", "unpatchify", "timestep_embedding", "cfg_scheduler_step")}
dense.linear = wrap("linear", orig["linear"], key=lambda a, k: (k.get("epilogue", 0), a[1].shape[0], a[1].shape[1]))
for n in ("attention", "layernorm_modulate", "linear_smallm", "patchify", "unpatchify", "timestep_embedding", "cfg_scheduler_step"):
    setattr(dense, n, wrap(n, orig[n]))
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record(); lat = pipe.denoise_step(lat, pe, int(ts[3]), 6.0, rope); s1.record(); torch.cuda.synchronize()
tot = s0.elapsed_time(s1)
print(f"step total {tot:.1f} ms")
acc = 0
for k, v in sorted(events.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1])):
    ms = sum(a.elapsed_time(b) for a, b in v); acc += ms
    print(f"  {k:40s} n={len(v):4d} total {ms:8.2f} ms  avg {ms/len(v):7.3f} ms  {100*ms/tot:5.1f}%")
print(f"  sum of measured {acc:.1f} ms; unattributed {tot-acc:.1f} ms")
