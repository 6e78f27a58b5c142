"""Dev timing of the full-size DPO training step (CogVideoX-5B shapes, BASELINE.json configs[4] per-GPU work: B = 1 pair) — dev tool."""
import sys, time, os
import torch
sys.path.insert(0, ".")
from videogpa_b200.train_dit import LoRATrainableTransformer
from videogpa_b200.train_step import DPOSharedStep
from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig

L = int(os.environ.get("LAYERS", "42"))
cfg = TransformerConfig.cogvideox_5b(); cfg.num_layers = L
base = CogVideoXTransformer3D.random_init(cfg, seed=1234, device="cuda")
ck = {"full": True, "mlp": "mlp", "none": False}[os.environ.get("CKPT", "full")]
pol = LoRATrainableTransformer(base, r=64, lora_alpha=128.0, gradient_checkpointing=ck)
print("checkpointing:", ck)
step = DPOSharedStep(base, None, beta=1.0, trainable=pol)
opt = step.configure_optimizers()
g = torch.Generator().manual_seed(0)
batch = {"x_win": torch.randn(1, 16, 13, 60, 90, generator=g), "x_lose": torch.randn(1, 16, 13, 60, 90, generator=g),
         "prompt_emb": torch.randn(1, 226, 4096, generator=g).to(torch.bfloat16)}
for it in range(int(os.environ.get("ITERS", "3"))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    opt.zero_grad(set_to_none=True)
    loss = step.training_step(batch)
    ev[1].record()
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record(); torch.cuda.synchronize()
    gn = sum(float(p.grad.float().pow(2).sum()) for p in pol.parameters()) ** 0.5
    print(f"step {it}: total {ev[0].elapsed_time(ev[3]):.0f} ms = forward (2 ref + 2 policy samples) {ev[0].elapsed_time(ev[1]):.0f} + backward (recompute + grads) "
          f"{ev[1].elapsed_time(ev[2]):.0f} + AdamW {ev[2].elapsed_time(ev[3]):.1f}; loss {loss.item():.5f} grad-norm {gn:.3e} "
          f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB, wall {time.perf_counter()-t0:.2f} s", flush=True)
