"""Dev: tiny invocation of every kernel family for compute-sanitizer (memcheck / racecheck) — not part of the product."""
import ctypes as C, math, sys
import torch
sys.path.insert(0, ".")
from videogpa_b200 import _lib, dense
from videogpa_b200.geometry import batch_reproject, get_colored_pointcloud, unproject_depth
from videogpa_b200.loss import DPOLoss
from videogpa_b200.metrics import MSEMetric, compute_motion_score_vectorized, epipolar_from_matches, mvcs_batch
from videogpa_b200.vae import AutoencoderKLCogVideoXDecoder, VAEDecoderConfig
from videogpa_b200.wan import WanConfig, WanTransformer3D
from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig
from videogpa_b200.rope import get_3d_rotary_pos_embed
BF = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
# DiT (2 blocks, small grid) incl. GEMM epilogues, attention d64, LN, embed, scheduler step
cfg = TransformerConfig(num_attention_heads=4, num_layers=2, text_embed_dim=256, sample_width=24, sample_height=16, sample_frames=9, max_text_seq_length=18)
m = CogVideoXTransformer3D.random_init(cfg, seed=1, device="cuda")
x = torch.randn(2, 3, 16, 16, 24, device="cuda", generator=g).to(BF)
e = torch.randn(2, 18, 256, device="cuda", generator=g).to(BF)
out = m(x, encoder_hidden_states=e, timestep=torch.tensor([999, 500], device="cuda"), image_rotary_emb=get_3d_rotary_pos_embed(64, 8, 12, 3), return_dict=False)[0]
dense.cfg_scheduler_step(out[1:2].contiguous(), out[0:1].contiguous(), x[:1].contiguous(), mode=dense.SCHED_DDIM, guidance=6.0, sqrt_alpha_t=0.5, sqrt_beta_t=0.8, c_sample=0.9, c_x0=0.1)
# Wan (1 block) incl. attention d128, rmsnorm+rope
w = WanTransformer3D.random_init(WanConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=128, text_len=64), seed=2, device="cuda")
w([torch.randn(48, 2, 8, 8, device="cuda", generator=g).to(BF)], torch.tensor([500.0]), [torch.randn(20, 128, device="cuda", generator=g).to(BF)])
# VAE decoder (small channels), tiled
dec = AutoencoderKLCogVideoXDecoder.random_init(VAEDecoderConfig(block_out_channels=(64, 64, 64, 128), sample_height=96, sample_width=160), seed=3, device="cuda")
dec.enable_tiling()
dec.decode(torch.randn(1, 16, 3, 12, 20, device="cuda", generator=g).to(BF))
# scorer
T, H, W = 4, 40, 56
depth = 2.0 + 0.5 * torch.rand(T, H, W, device="cuda", generator=g)
K = torch.tensor([[0.8 * W, 0, W / 2], [0, 0.8 * W, H / 2], [0, 0, 1]], device="cuda").expand(T, 3, 3).contiguous()
E = torch.zeros(T, 3, 4, device="cuda"); E[:, :3, :3] = torch.eye(3, device="cuda"); E[:, 0, 3] = 0.02 * torch.arange(T, device="cuda")
mvcs_batch(depth[None], K[None], E[None])
world = unproject_depth(depth, K, E)
imgs = torch.rand(T, 3, H, W, device="cuda", generator=g)
v, c = get_colored_pointcloud(dict(world_points_from_depth=world, depth_conf=1 + torch.rand(T, H, W, device="cuda", generator=g), images=imgs), mode="depth", conf_thres=30)
rep = batch_reproject(v, c, K, E, H, W)
MSEMetric().compute(gt=imgs, rep=rep); compute_motion_score_vectorized(E)
p1 = torch.rand(2, 64, 2, device="cuda", generator=g) * 100
epipolar_from_matches(p1, p1 + torch.rand(2, 64, 2, device="cuda", generator=g), None)
ts = [torch.randn(2, 3, 4, 8, 8, device="cuda", generator=g) for _ in range(6)]
ts[0].requires_grad_(True); ts[1].requires_grad_(True)
DPOLoss(beta=2.0)(*ts).loss.backward()
# round-1 additions: VAE encoder (plain GroupNorm path, strided conv via subsampling), T5 encoder (bias attention, gated product),
# training step (attention backward, row-wise backward kernels, LoRA dgrad / wgrad GEMMs) with an odd token count
from videogpa_b200.vae import AutoencoderKLCogVideoXEncoder
from videogpa_b200.t5 import T5Config, T5EncoderModel
from videogpa_b200.train_dit import LoRATrainableTransformer
from videogpa_b200.train_step import DPOSharedStep
enc = AutoencoderKLCogVideoXEncoder.random_init(VAEDecoderConfig(block_out_channels=(64, 64, 64, 128), sample_height=96, sample_width=160), seed=4, device="cuda")
enc.enable_tiling()
enc.encode((torch.rand(1, 3, 9, 96, 160, device="cuda", generator=g) * 2 - 1).to(BF))
t5 = T5EncoderModel.random_init(T5Config(vocab_size=64, d_model=256, d_ff=512, num_layers=1, num_heads=4), seed=5, device="cuda")
t5.use_cuda_graph = False
t5(torch.randint(0, 64, (2, 37), device="cuda", generator=g))
pol = LoRATrainableTransformer(m, r=64, lora_alpha=128.0, gradient_checkpointing="mlp")
with torch.no_grad():
    for layer in pol.lora:
        for k in layer:
            layer[k][1].normal_(0, 0.02)
step = DPOSharedStep(m, None, beta=5.0, trainable=pol)
gc = torch.Generator().manual_seed(1)
batch = {"x_win": torch.randn(1, 16, 3, 16, 24, generator=gc), "x_lose": torch.randn(1, 16, 3, 16, 24, generator=gc),
         "prompt_emb": torch.randn(1, 18, 256, generator=gc).to(BF)}          # 2 x 306 tokens: not a multiple of 8, partial tiles everywhere
step.training_step(batch).backward()
# round-2 additions: exact attention kernel next to the bounded one (and a call whose heads split between them), the three-tier MVCS
# kernel on an odd width, both T5 attention kernels (tensor-core path S <= 256, CUDA-core path above), 128-wide skinny GEMM tiles,
# the re-scalable LoRA attachment, the temporal-patch (CogVideoX1.5) training forward + backward
qq = torch.randn(1, 300, 128, device="cuda", generator=g).to(BF); kk = torch.randn(1, 261, 128, device="cuda", generator=g).to(BF)
vv = torch.randn(1, 261, 128, device="cuda", generator=g).to(BF)
dense.attention(qq, kk, vv, 2, exact=True)
qq2 = qq.clone(); qq2[..., 64:] *= 40.0
dense.attention(qq2, kk, vv, 2)
depth_o = 2.0 + 0.5 * torch.rand(2, 3, 37, 57, device="cuda", generator=g)
mvcs_batch(depth_o, K[None, :3].expand(2, 3, 3, 3).contiguous(), E[None, :3].expand(2, 3, 3, 4).contiguous())
t5(torch.randint(0, 64, (1, 300), device="cuda", generator=g))
dense.linear(torch.randn(226, 256, device="cuda", generator=g).to(BF), (torch.randn(128 * 80, 256, device="cuda", generator=g) * 0.05).to(BF), None)
cfg15 = TransformerConfig(num_attention_heads=4, num_layers=1, text_embed_dim=256, sample_width=24, sample_height=16, sample_frames=9,
                          max_text_seq_length=18, patch_size_t=2)
m15 = CogVideoXTransformer3D.random_init(cfg15, seed=6, device="cuda")
pol15 = LoRATrainableTransformer(m15, r=64, lora_alpha=128.0, gradient_checkpointing=True)
step15 = DPOSharedStep(m15, None, beta=1.0, trainable=pol15)
batch15 = {"x_win": torch.randn(1, 16, 5, 17, 25, generator=gc), "x_lose": torch.randn(1, 16, 5, 17, 25, generator=gc),
           "prompt_emb": torch.randn(1, 18, 256, generator=gc).to(BF)}
step15.training_step(batch15).backward()
torch.cuda.synchronize()
print("sanitize run complete")
