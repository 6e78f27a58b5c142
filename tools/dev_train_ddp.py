"""DDP check of the DPO training step on N GPUs (torchrun): every rank trains on its own preference pair, LoRA gradients are
all-reduced over NCCL (parallel.average_gradients), parameters must stay bit-identical across ranks. Prints the step time
(max over ranks). Dev tool: LAYERS env var shrinks the model."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from videogpa_b200.parallel import init_from_env
from videogpa_b200.train_dit import LoRATrainableTransformer
from videogpa_b200.train_step import DPOSharedStep
from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig

rank, world, local = init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cfg = TransformerConfig.cogvideox_5b(); cfg.num_layers = int(os.environ.get("LAYERS", "42"))
base = CogVideoXTransformer3D.random_init(cfg, seed=1234, device=dev)
pol = LoRATrainableTransformer(base, r=64, lora_alpha=128.0, seed=0)
with torch.no_grad():                                    # non-zero B so that every factor gets a gradient
    g0 = torch.Generator(device=dev).manual_seed(99)
    for layer in pol.lora:
        for m in layer:
            layer[m][1].copy_(0.01 * torch.randn(layer[m][1].shape, device=dev, generator=g0))
step = DPOSharedStep(base, None, beta=1.0, trainable=pol)
opt = step.configure_optimizers(lr=1e-4)
g = torch.Generator().manual_seed(100 + rank)            # a different preference pair per rank
batch = {"x_win": torch.randn(1, 16, 13, 60, 90, generator=g), "x_lose": torch.randn(1, 16, 13, 60, 90, generator=g),
         "prompt_emb": torch.randn(1, 226, 4096, generator=g).to(torch.bfloat16)}
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = step.fit_step(batch, opt)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    chk = torch.stack([p.detach().double().sum() for p in pol.parameters()]).sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(torch.equal(allc[0], c) for c in allc)
    if rank == 0:
        print(f"step {it}: {ms.item():.0f} ms (max over {world} ranks), {world * 1000.0 / ms.item():.3f} pairs/s, loss[rank0] {loss:.5f}, "
              f"parameters identical across ranks: {same}", flush=True)
    assert same
dist.barrier(); dist.destroy_process_group()
