"""Dev: one launch of every kernel family at its production shape, for `ncu --set full` (not part of the product)."""
import sys, math
import torch
sys.path.insert(0, ".")
from videogpa_b200 import dense
from videogpa_b200.metrics import mvcs_batch
from videogpa_b200.vae import AutoencoderKLCogVideoXDecoder, VAEDecoderConfig
from videogpa_b200.geometry import batch_reproject
BF = torch.bfloat16
M, S, St, D = 35552, 17776, 226, 3072
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, D, device="cuda", generator=g).to(BF)
n = torch.empty_like(x)
w_qkv = (torch.randn(3 * D, D, device="cuda", generator=g) * 0.02).to(BF); b_qkv = torch.zeros(3 * D, device="cuda", dtype=BF)
w_o = (torch.randn(D, D, device="cuda", generator=g) * 0.02).to(BF); b_o = torch.zeros(D, device="cuda", dtype=BF)
w_f1 = (torch.randn(4 * D, D, device="cuda", generator=g) * 0.02).to(BF); b_f1 = torch.zeros(4 * D, device="cuda", dtype=BF)
mod = torch.randn(2, 6 * D, device="cuda", generator=g).to(BF)
ones, zeros = torch.ones(D, device="cuda", dtype=BF), torch.zeros(D, device="cuda", dtype=BF)
lnq = (torch.ones(64, device="cuda"), torch.zeros(64, device="cuda"))
ang = torch.rand(S - St, 32, device="cuda") * 6.28
rope = (ang.cos().repeat_interleave(2, 1).contiguous(), ang.sin().repeat_interleave(2, 1).contiguous())
seg = dict(rows_per_sample=S, text_rows=St)
qkv = torch.empty(M, 3 * D, device="cuda", dtype=BF)
ffh = torch.empty(M, 4 * D, device="cuda", dtype=BF)
for _ in range(2):
    dense.layernorm_modulate(x, ones, zeros, eps=1e-5, out=n, **seg, shift_vid=mod[:, 0:D], scale_vid=mod[:, D:2 * D], shift_txt=mod[:, 3 * D:4 * D],
                             scale_txt=mod[:, 4 * D:5 * D], mod_stride_b=6 * D)
    dense.linear(n, w_qkv, b_qkv, out=qkv, epilogue=dense.EPI_QKV, **seg, ln_q=lnq, ln_k=lnq, ln_eps=1e-6, rope=rope, model_dim=D)
    dense.linear(n, w_o, b_o, out=x, epilogue=dense.EPI_GATE_RES, **seg, gate_vid=mod[:, 2 * D:3 * D], gate_txt=mod[:, 5 * D:6 * D], gate_stride_b=6 * D)
    dense.linear(n, w_f1, b_f1, out=ffh, epilogue=dense.EPI_BIAS_GELU)
# scorer
N, T, H, W = 32, 10, 504, 504
depth = 2.0 + 0.5 * torch.rand(N, T, H, W, device="cuda", generator=g)
K = torch.tensor([[0.8 * W, 0, W / 2], [0, 0.8 * W, H / 2], [0, 0, 1]], device="cuda").expand(N, T, 3, 3).contiguous()
E = torch.zeros(N, T, 3, 4, device="cuda")
for i in range(T):
    a = math.radians(0.5 * i)
    E[:, i] = torch.tensor([[math.cos(a), 0, math.sin(a), 0.02 * i], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0]], device="cuda")
for _ in range(2):
    mvcs_batch(depth, K, E)
pts = torch.randn(T * H * W, 3, device="cuda", generator=g) * 0.5 + torch.tensor([0, 0, 3.0], device="cuda")
cols = torch.rand(T * H * W, 3, device="cuda", generator=g) * 255
for _ in range(2):
    batch_reproject(pts, cols, K[0], E[0], H, W)
# VAE: one tile, one frame batch at the real channel counts (30x45 latent tile, 2 latent frames)
if "vae" not in sys.argv:
    torch.cuda.synchronize(); print("done (no vae)"); sys.exit(0)
dec = AutoencoderKLCogVideoXDecoder.random_init(VAEDecoderConfig(), seed=5, device="cuda")
z = torch.randn(1, 16, 2, 30, 45, device="cuda", generator=g).to(BF)
dec.decode(z)
torch.cuda.synchronize()
print("done")
