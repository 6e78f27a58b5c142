"""Dev check of the DiT kernels against torch on the GPU box (not part of the product)."""
import sys, math
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from videogpa_b200 import dense

torch.manual_seed(0)
dev = "cuda"
BF = torch.bfloat16

def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()

def check_attn(B, H, S, Skv=None):
    Skv = Skv or S
    D = H * 64
    qkv = (torch.randn(B, S, 3 * D, device=dev) * 1.0).to(BF)
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    if Skv != S:
        kv = torch.randn(B, Skv, 2 * D, device=dev).to(BF)
        k, v = kv[..., :D], kv[..., D:]
    out = dense.attention(q, k, v, H)
    torch.cuda.synchronize()
    qh = q.reshape(B, S, H, 64).transpose(1, 2).float()
    kh = k.reshape(B, Skv, H, 64).transpose(1, 2).float()
    vh = v.reshape(B, Skv, H, 64).transpose(1, 2).float()
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, S, D)
    print(f"attn B={B} H={H} S={S} Skv={Skv}: rel err {rel(out, ref):.3e} finite={torch.isfinite(out.float()).all().item()}", flush=True)

for (B, H, S) in [(1, 1, 128), (1, 2, 256), (1, 2, 300), (2, 3, 1000), (1, 2, 4096)]:
    check_attn(B, H, S)
check_attn(1, 2, 300, 517)
check_attn(1, 1, 17776)

def bench_attn(B, H, S, iters=5):
    D = H * 64
    qkv = torch.randn(B, S, 3 * D, device=dev).to(BF)
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    out = torch.empty(B, S, D, device=dev, dtype=BF)
    for _ in range(2):
        dense.attention(q, k, v, H, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dense.attention(q, k, v, H, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * B * H * S * S * 64
    qh = q.reshape(B, S, H, 64).transpose(1, 2)
    kh = k.reshape(B, S, H, 64).transpose(1, 2)
    vh = v.reshape(B, S, H, 64).transpose(1, 2)
    for _ in range(2):
        F.scaled_dot_product_attention(qh, kh, vh)
    e0.record()
    for _ in range(iters):
        F.scaled_dot_product_attention(qh, kh, vh)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"bench attn B={B} H={H} S={S}: {ms:.3f} ms {fl/ms/1e9:.1f} TF/s | torch sdpa {ms2:.3f} ms {fl/ms2/1e9:.1f} TF/s", flush=True)

bench_attn(2, 48, 17776)

# ---------------------------------------------------------------- fused QKV epilogue
def check_qkv(M, D, text_rows, rows_per_sample, use_rope=True):
    x = (torch.randn(M, D, device=dev) * 0.5).to(BF)
    w = (torch.randn(3 * D, D, device=dev) * 0.03).to(BF)
    b = (torch.randn(3 * D, device=dev) * 0.1).to(BF)
    lqw, lqb = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
    lkw, lkb = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
    nv = rows_per_sample - text_rows
    ang = torch.rand(nv, 32, device=dev) * 6.28
    cos = torch.cos(ang).repeat_interleave(2, 1).contiguous()
    sin = torch.sin(ang).repeat_interleave(2, 1).contiguous()
    out = dense.linear(x, w, b, epilogue=dense.EPI_QKV, rows_per_sample=rows_per_sample, text_rows=text_rows,
                       ln_q=(lqw, lqb), ln_k=(lkw, lkb), ln_eps=1e-6, rope=(cos, sin) if use_rope else None, model_dim=D)
    torch.cuda.synchronize()
    y = (x.float() @ w.float().t() + b.float()).to(BF)
    H = D // 64
    Bn = M // rows_per_sample
    def norm_rope(t, lw, lb):
        t = t.reshape(Bn, rows_per_sample, H, 64).float()
        t = F.layer_norm(t, (64,), lw, lb, 1e-6).to(BF)
        if use_rope:
            tv = t[:, text_rows:].float()
            xr, xi = tv.reshape(*tv.shape[:-1], 32, 2).unbind(-1)
            rot = torch.stack([-xi, xr], -1).flatten(-2)
            tv = (tv * cos[None, :, None, :] + rot * sin[None, :, None, :]).to(BF)
            t = torch.cat([t[:, :text_rows], tv], 1)
        return t.reshape(M, D)
    ref = torch.cat([norm_rope(y[:, :D], lqw, lqb), norm_rope(y[:, D:2 * D], lkw, lkb), y[:, 2 * D:]], 1)
    print(f"qkv epilogue M={M} D={D} rope={use_rope}: rel err {rel(out, ref):.3e}", flush=True)

check_qkv(2 * 500, 256, 26, 500)
check_qkv(2 * 500, 256, 26, 500, use_rope=False)

# ---------------------------------------------------------------- gate residual / accum
def check_gate(M, N, K, text_rows, rows_per_sample):
    x = (torch.randn(M, K, device=dev) * 0.5).to(BF)
    w = (torch.randn(N, K, device=dev) * 0.03).to(BF)
    b = (torch.randn(N, device=dev) * 0.1).to(BF)
    Bn = M // rows_per_sample
    res = torch.randn(M, N, device=dev).to(BF)
    gates = torch.randn(Bn, 2 * N, device=dev).to(BF)
    out = res.clone()
    dense.linear(x, w, b, out=out, epilogue=dense.EPI_GATE_RES, rows_per_sample=rows_per_sample, text_rows=text_rows,
                 gate_txt=gates[:, :N], gate_vid=gates[:, N:], gate_stride_b=gates.stride(0))
    torch.cuda.synchronize()
    y = (x.float() @ w.float().t() + b.float()).to(BF).reshape(Bn, rows_per_sample, N)
    g = torch.where((torch.arange(rows_per_sample, device=dev) < text_rows)[None, :, None], gates[:, None, :N], gates[:, None, N:])
    ref = (res.reshape(Bn, rows_per_sample, N) + (g * y)).reshape(M, N)
    print(f"gate_res M={M} N={N} K={K}: rel err {rel(out, ref):.3e}", flush=True)

check_gate(2 * 500, 256, 512, 26, 500)

def check_accum():
    r = 64; D = 512
    A = torch.randn(r, D, device=dev) * 0.05; Bm = torch.randn(D, r, device=dev) * 0.05
    W = (torch.randn(D, D, device=dev) * 0.03).to(BF)
    ref = (W.float() + 2.0 * (Bm @ A)).to(BF)
    def split(t):
        hi = t.to(BF); lo = (t - hi.float()).to(BF); return hi, lo
    Bh, Bl = split(Bm); Ah, Al = split(A.t().contiguous())
    a_op = torch.cat([Bh, Bh, Bl], 1).contiguous()      # [D, 3r]
    w_op = torch.cat([Ah, Al, Ah], 1).contiguous()      # [D(in), 3r] -> N = in features
    out = W.clone()
    dense.linear(a_op, w_op, None, out=out, epilogue=dense.EPI_ACCUM, alpha=2.0)
    torch.cuda.synchronize()
    print(f"lora accum: mismatching bf16 elements {(out != ref).sum().item()} / {out.numel()}, rel {rel(out, ref):.3e}", flush=True)

check_accum()

# ---------------------------------------------------------------- layernorm modulate
def check_ln(rows, D, text_rows, rows_per_sample):
    x = torch.randn(rows, D, device=dev).to(BF)
    w = (torch.rand(D, device=dev) + 0.5).to(BF); b = (torch.randn(D, device=dev) * 0.1).to(BF)
    Bn = rows // rows_per_sample
    mod = (torch.randn(Bn, 6 * D, device=dev) * 0.3).to(BF)
    out = dense.layernorm_modulate(x, w, b, eps=1e-5, rows_per_sample=rows_per_sample, text_rows=text_rows,
                                   shift_vid=mod[:, 0:D], scale_vid=mod[:, D:2 * D], shift_txt=mod[:, 3 * D:4 * D],
                                   scale_txt=mod[:, 4 * D:5 * D], mod_stride_b=mod.stride(0))
    torch.cuda.synchronize()
    n = F.layer_norm(x, (D,), w, b, 1e-5).reshape(Bn, rows_per_sample, D)
    is_t = (torch.arange(rows_per_sample, device=dev) < text_rows)[None, :, None]
    sc = torch.where(is_t, mod[:, None, 4 * D:5 * D], mod[:, None, D:2 * D])
    sh = torch.where(is_t, mod[:, None, 3 * D:4 * D], mod[:, None, 0:D])
    ref = (n * (1 + sc) + sh).reshape(rows, D)
    print(f"ln_modulate rows={rows} D={D}: rel err {rel(out, ref):.3e} exact-mismatch {(out != ref).float().mean().item():.2e}", flush=True)
    out2 = dense.layernorm_modulate(x, w, b, eps=1e-5)
    print(f"ln plain: rel err {rel(out2, F.layer_norm(x, (D,), w, b, 1e-5)):.3e}", flush=True)

check_ln(2 * 500, 3072, 26, 500)

def bench_ln():
    rows, D = 35552, 3072
    x = torch.randn(rows, D, device=dev).to(BF)
    w = torch.ones(D, device=dev).to(BF); b = torch.zeros(D, device=dev).to(BF)
    mod = torch.randn(2, 6 * D, device=dev).to(BF)
    out = torch.empty_like(x)
    kw = dict(eps=1e-5, out=out, rows_per_sample=17776, text_rows=226, shift_vid=mod[:, 0:D], scale_vid=mod[:, D:2 * D],
              shift_txt=mod[:, 3 * D:4 * D], scale_txt=mod[:, 4 * D:5 * D], mod_stride_b=mod.stride(0))
    for _ in range(3): dense.layernorm_modulate(x, w, b, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dense.layernorm_modulate(x, w, b, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"bench ln_modulate: {ms:.3f} ms, {2 * rows * D * 2 / ms / 1e6:.0f} GB/s", flush=True)
bench_ln()

# ---------------------------------------------------------------- conditioning path
def check_small():
    x = torch.randn(2, 512, device=dev).to(BF)
    w = (torch.randn(18432, 512, device=dev) * 0.03).to(BF); b = (torch.randn(18432, device=dev) * 0.1).to(BF)
    out = dense.linear_smallm(x, w, b, act_in=dense.ACT_SILU)
    ref = F.linear(F.silu(x), w, b)
    print(f"smallm silu: rel err {rel(out, ref):.3e}", flush=True)
    t = torch.tensor([999.0, 19.0], device=dev)
    te = dense.timestep_embedding(t, 3072)
    half = 1536
    fr = torch.exp(-math.log(10000) * torch.arange(half, device=dev, dtype=torch.float32) / half)
    e = t[:, None] * fr[None]
    ref = torch.cat([torch.cos(e), torch.sin(e)], -1).to(BF)
    print(f"timestep emb: max abs err {(te.float() - ref.float()).abs().max().item():.3e}", flush=True)
    lat = torch.randn(4, 16, 60, 90, device=dev).to(BF)
    pt = dense.patchify(lat)
    ref = lat.reshape(4, 16, 30, 2, 45, 2).permute(0, 2, 4, 1, 3, 5).reshape(4 * 1350, 64)
    print(f"patchify exact: {torch.equal(pt, ref)}", flush=True)
    back = dense.unpatchify(pt, 4, 16, 60, 90)
    print(f"unpatchify roundtrip exact: {torch.equal(back, lat)}", flush=True)
check_small()
