"""GPU parity of the training-step kernels (backward of the DiT for the DPO step) against torch
autograd in fp32 on bf16-rounded inputs. Tolerances: attention gradients <= 2e-2 of the max (P and dS are rounded to bf16
before the second GEMMs, as in every flash-attention backward), row-wise ops <= 1e-2."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def relmax(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


def _attn_ref(q, k, v, d_out, H, scale):
    B, S, _ = q.shape
    qf, kf, vf = [t.float().view(B, -1, H, 64).transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v)]
    s = (qf @ kf.transpose(-1, -2)) * scale
    o = torch.softmax(s, dim=-1) @ vf
    o.backward(d_out.float().view(B, -1, H, 64).transpose(1, 2))
    lse2 = torch.logsumexp(s, dim=-1) * math.log2(math.e)
    back = lambda t: t.transpose(1, 2).reshape(B, -1, H * 64)
    return back(o.detach()), lse2.detach(), back(qf.grad), back(kf.grad), back(vf.grad)


@pytest.mark.parametrize("B,H,Sq,Skv", [(1, 2, 128, 128), (2, 3, 300, 300), (1, 2, 200, 450), (1, 1, 40, 40), (1, 2, 64, 65), (1, 1, 1, 193), (1, 1, 385, 8)])
def test_attention_backward_vs_autograd(lib, B, H, Sq, Skv):
    from videogpa_b200 import dense
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(B, Sq, H * 64, generator=g) * 1.5).to(BF)
    k = (torch.randn(B, Skv, H * 64, generator=g) * 1.5).to(BF)
    v = torch.randn(B, Skv, H * 64, generator=g).to(BF)
    d_out = torch.randn(B, Sq, H * 64, generator=g).to(BF)
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = _attn_ref(q, k, v, d_out, H, 0.125)
    qc, kc, vc, dc = q.cuda(), k.cuda(), v.cuda(), d_out.cuda()
    lse = torch.empty(B, H, Sq, dtype=torch.float32, device="cuda")
    out = dense.attention(qc, kc, vc, H, lse=lse)
    assert relmax(out.cpu(), o_ref) < 1e-2
    assert (lse.cpu() - lse_ref).abs().max().item() < 2e-2          # log2 units
    dq, dk, dv = dense.attention_backward(qc, kc, vc, out, dc, lse, H)
    torch.cuda.synchronize()
    assert relmax(dq.cpu(), dq_ref) < 2e-2, ("dq", relmax(dq.cpu(), dq_ref))
    assert relmax(dk.cpu(), dk_ref) < 2e-2, ("dk", relmax(dk.cpu(), dk_ref))
    assert relmax(dv.cpu(), dv_ref) < 2e-2, ("dv", relmax(dv.cpu(), dv_ref))


def test_attention_backward_fused_qkv_views(lib):
    """q / k / v as column slices of one fused projection buffer (row stride 3*H*64), as the DiT block holds them."""
    from videogpa_b200 import dense
    g = torch.Generator().manual_seed(6)
    B, H, S = 1, 2, 260
    qkv = torch.randn(B, S, 3 * H * 64, generator=g).to(BF)
    d_out = torch.randn(B, S, H * 64, generator=g).to(BF)
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    _, _, dq_ref, dk_ref, dv_ref = _attn_ref(q, k, v, d_out, H, 0.125)
    dev = qkv.cuda()
    qc, kc, vc = dev[..., :H * 64], dev[..., H * 64:2 * H * 64], dev[..., 2 * H * 64:]
    lse = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
    out = dense.attention(qc, kc, vc, H, lse=lse)
    dq, dk, dv = dense.attention_backward(qc, kc, vc, out, d_out.cuda(), lse, H)
    for got, ref in ((dq, dq_ref), (dk, dk_ref), (dv, dv_ref)):
        assert relmax(got.cpu(), ref) < 2e-2


# ---------------------------------------------------------------------------------------------- row-wise backward kernels
def test_layernorm_modulate_backward_vs_autograd(lib):
    import ctypes as C
    from videogpa_b200 import _lib
    from videogpa_b200.train_dit import _ln_args
    L = _lib.load()
    g = torch.Generator().manual_seed(1)
    B, S, St, D = 2, 37, 5, 512
    x = torch.randn(B * S, D, generator=g).to(BF)
    w = (1 + 0.2 * torch.randn(D, generator=g)).to(BF)
    bias = (0.1 * torch.randn(D, generator=g)).to(BF)
    mod = (0.5 * torch.randn(B, 4 * D, generator=g)).to(BF)          # per sample: scale_txt | scale_vid | (unused shifts)
    dy = torch.randn(B * S, D, generator=g).to(BF)
    add = torch.randn(B * S, D, generator=g).to(BF)
    xf = x.float().requires_grad_(True)
    scale = torch.empty(B * S, D)
    for b in range(B):
        scale[b * S:b * S + St] = mod[b, :D].float()
        scale[b * S + St:(b + 1) * S] = mod[b, D:2 * D].float()
    y = torch.nn.functional.layer_norm(xf, (D,), w.float(), bias.float(), 1e-5) * (1 + scale)
    y.backward(dy.float())
    ref = xf.grad + add.float()
    xc, dyc, addc, modc, wc = x.cuda(), dy.cuda(), add.cuda(), mod.cuda(), w.cuda()
    dx = torch.empty_like(xc)
    a = _ln_args(xc, wc, 1e-5, (S, St), modc[:, :D], modc[:, D:2 * D], 4 * D)
    _lib.check(L.vgpa_layernorm_modulate_bwd_bf16(C.byref(a), dyc.data_ptr(), D, addc.data_ptr(), D, dx.data_ptr(), D, _lib.current_stream()), "ln bwd")
    assert relmax(dx.cpu(), ref) < 1e-2
    # plain LayerNorm without affine / modulation / add
    xf2 = x.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xf2, (D,), None, None, 1e-5).backward(dy.float())
    a2 = _ln_args(xc, None, 1e-5, (0, 0), None, None, 0)
    _lib.check(L.vgpa_layernorm_modulate_bwd_bf16(C.byref(a2), dyc.data_ptr(), D, None, 0, dx.data_ptr(), D, _lib.current_stream()), "ln bwd")
    assert relmax(dx.cpu(), xf2.grad) < 1e-2


def test_head_layernorm_gelu_gate_kernels_vs_autograd(lib):
    from videogpa_b200 import _lib
    from videogpa_b200.train_dit import _Gelu, _GateRes, _HeadLN
    g = torch.Generator().manual_seed(2)
    M, heads = 50, 3
    D = heads * 64
    qkv = torch.randn(M, 3 * D, generator=g).to(BF)
    lnq = ((1 + 0.2 * torch.randn(64, generator=g)), 0.1 * torch.randn(64, generator=g))
    lnk = ((1 + 0.2 * torch.randn(64, generator=g)), 0.1 * torch.randn(64, generator=g))
    dy = torch.randn(M, 3 * D, generator=g).to(BF)
    xf = qkv.float().requires_grad_(True)
    parts = []
    for i, (w, b) in enumerate((lnq, lnk)):
        parts.append(torch.nn.functional.layer_norm(xf[:, i * D:(i + 1) * D].reshape(M, heads, 64), (64,), w, b, 1e-6).reshape(M, D))
    ref = torch.cat(parts + [xf[:, 2 * D:]], dim=1)
    ref.backward(dy.float())
    xc = qkv.cuda().requires_grad_(True)
    out = _HeadLN.apply(xc, heads, tuple(t.cuda() for t in lnq), tuple(t.cuda() for t in lnk), 1e-6)
    out.backward(dy.cuda())
    assert relmax(out.detach().cpu(), ref.detach()) < 1e-2 and relmax(xc.grad.cpu(), xf.grad) < 1e-2
    # GELU(tanh)
    x = (2 * torch.randn(40, 64, generator=g)).to(BF)
    d = torch.randn(40, 64, generator=g).to(BF)
    xf = x.float().requires_grad_(True)
    r = torch.nn.functional.gelu(xf, approximate="tanh")
    r.backward(d.float())
    xc = x.cuda().requires_grad_(True)
    o = _Gelu.apply(xc)
    o.backward(d.cuda())
    assert relmax(o.detach().cpu(), r.detach()) < 1e-2 and relmax(xc.grad.cpu(), xf.grad) < 1e-2
    # gated residual: hidden + gate[b, seg] * branch
    B, S, St, D = 2, 11, 3, 64
    hid, br = torch.randn(B * S, D, generator=g).to(BF), torch.randn(B * S, D, generator=g).to(BF)
    gates = torch.randn(B, 2 * D, generator=g).to(BF)
    gfull = torch.empty(B * S, D)
    for b in range(B):
        gfull[b * S:b * S + St] = gates[b, :D].float()
        gfull[b * S + St:(b + 1) * S] = gates[b, D:].float()
    d = torch.randn(B * S, D, generator=g).to(BF)
    hc, bc, gc = hid.cuda().requires_grad_(True), br.cuda().requires_grad_(True), gates.cuda()
    o = _GateRes.apply(hc, bc, (S, St), gc[:, :D], gc[:, D:], 2 * D)
    o.backward(d.cuda())
    want = (hid.float() + (gfull * br.float()).to(BF).float()).to(BF)
    assert torch.equal(o.detach().cpu(), want)                                     # bf16 product, then bf16 sum: bit-exact
    assert torch.equal(hc.grad.cpu(), d) and torch.equal(bc.grad.cpu(), (d.float() * gfull).to(BF))


# ---------------------------------------------------------------------------------------------- the whole training step
def _torch_dpo_loss(vw, vl, rw, rl, tw, tl, beta):
    err = lambda a, b: ((a - b) ** 2).reshape(a.shape[0], -1).mean(dim=1)
    logits = beta * ((err(rw, tw) - err(vw, tw)) - (err(rl, tl) - err(vl, tl)))
    return torch.nn.functional.softplus(-logits).mean()


def test_dpo_training_step_gradients_vs_oracle_autograd(lib):
    """training_step + backward() against torch autograd through the oracle DiT (fp32, CPU) with the LoRA delta merged into the
    policy weights (the same function of A, B as PEFT's unmerged branch): loss and every dA / dB within 6e-2 of the tensor max
    (two blocks of bf16 forward + backward; most tensors agree to ~1e-2)."""
    from oracle import dit_torch as O
    from videogpa_b200.train_dit import TARGETS, LoRATrainableTransformer
    from videogpa_b200.train_step import DPOSharedStep
    from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig
    kw = dict(num_attention_heads=4, num_layers=2, text_embed_dim=256, sample_width=24, sample_height=16, sample_frames=9, max_text_seq_length=18)
    ocfg = O.DiTConfig(**kw)
    sd = {k: v.to(BF).float() for k, v in O.random_state_dict(ocfg, seed=7, randomize_norms=True, std=0.05).items()}
    base = CogVideoXTransformer3D(TransformerConfig(**kw), sd, device="cuda")
    pol = LoRATrainableTransformer(base, r=64, lora_alpha=128.0, seed=3)
    g = torch.Generator().manual_seed(31)
    D = 256
    ref_params = {}
    for layer in range(2):
        for m in TARGETS:
            A = (0.05 * torch.randn(64, D, generator=g)).to(BF).float()
            Bm = (0.05 * torch.randn(D, 64, generator=g)).to(BF).float()
            with torch.no_grad():
                pol.lora[layer][m][0].copy_(A)
                pol.lora[layer][m][1].copy_(Bm)
            ref_params[(layer, m)] = (A.clone().requires_grad_(True), Bm.clone().requires_grad_(True))
    Bsz, C, Fr, H, W, St = 2, 16, 3, 16, 24, 18
    batch = {"x_win": torch.randn(Bsz, C, Fr, H, W, generator=g), "x_lose": torch.randn(Bsz, C, Fr, H, W, generator=g),
             "prompt_emb": torch.randn(Bsz, St, 256, generator=g).to(BF)}
    t = torch.tensor([812, 77])
    noise = torch.randn(Bsz, Fr, C, H, W, generator=g)
    beta = 50.0                                                    # a logit of O(1) so that the sigmoid factor matters
    step = DPOSharedStep(base, None, beta=beta, trainable=pol)
    loss = step.training_step(batch, timesteps=t.cuda(), noise=noise.cuda())
    loss.backward()
    torch.cuda.synchronize()

    # oracle: merged policy weights as differentiable functions of A, B
    sd_pol = dict(sd)
    for (layer, m), (A, Bm) in ref_params.items():
        k = f"transformer_blocks.{layer}.attn1.{m}.weight"
        sd_pol[k] = sd[k] + 2.0 * (Bm @ A)
    ac = O.cogvideox_alphas_cumprod()
    xw, xl = batch["x_win"].permute(0, 2, 1, 3, 4), batch["x_lose"].permute(0, 2, 1, 3, 4)
    fw = lambda s_, x: O.transformer_forward(s_, ocfg, O.add_noise(ac, x, noise, t.numpy()), batch["prompt_emb"].float(), t, None)
    with torch.no_grad():
        rw, rl = fw(sd, xw), fw(sd, xl)
    want = _torch_dpo_loss(fw(sd_pol, xw), fw(sd_pol, xl), rw, rl, O.get_velocity(ac, xw, noise, t.numpy()),
                           O.get_velocity(ac, xl, noise, t.numpy()), beta)
    want.backward()
    assert abs(loss.item() - want.item()) < 5e-2, (loss.item(), want.item())
    worst = 0.0
    for (layer, m), (A, Bm) in ref_params.items():
        gA, gB = pol.lora[layer][m][0].grad, pol.lora[layer][m][1].grad
        assert gA is not None and gB is not None and gA.dtype == torch.float32
        ea, eb = relmax(gA.cpu(), A.grad), relmax(gB.cpu(), Bm.grad)
        worst = max(worst, ea, eb)
        assert ea < 6e-2 and eb < 6e-2, (layer, m, ea, eb)
    # gradient checkpointing must not change the result
    pol2 = LoRATrainableTransformer(base, r=64, lora_alpha=128.0, seed=3, gradient_checkpointing=False)
    for layer in range(2):
        for m in TARGETS:
            with torch.no_grad():
                pol2.lora[layer][m][0].copy_(pol.lora[layer][m][0])
                pol2.lora[layer][m][1].copy_(pol.lora[layer][m][1])
    step2 = DPOSharedStep(base, None, beta=beta, trainable=pol2)
    step2.training_step(batch, timesteps=t.cuda(), noise=noise.cuda()).backward()
    for layer in range(2):
        for m in TARGETS:
            assert torch.equal(pol2.lora[layer][m][0].grad, pol.lora[layer][m][0].grad)
    # ... nor does recomputing only the MLP half
    pol3 = LoRATrainableTransformer(base, r=64, lora_alpha=128.0, seed=3, gradient_checkpointing="mlp")
    for layer in range(2):
        for m in TARGETS:
            with torch.no_grad():
                pol3.lora[layer][m][0].copy_(pol.lora[layer][m][0])
                pol3.lora[layer][m][1].copy_(pol.lora[layer][m][1])
    DPOSharedStep(base, None, beta=beta, trainable=pol3).training_step(batch, timesteps=t.cuda(), noise=noise.cuda()).backward()
    for layer in range(2):
        for m in TARGETS:
            assert torch.equal(pol3.lora[layer][m][1].grad, pol.lora[layer][m][1].grad)
    # one optimizer step (AdamW, lr 5e-6 as in 03_train.py:207-210) moves every factor
    opt = step.configure_optimizers()
    before = pol.lora[1]["to_q"][0].detach().clone()
    opt.step()
    assert not torch.equal(before, pol.lora[1]["to_q"][0].detach())
    names = [n for n, _ in pol.named_parameters()]
    assert names[0] == "base_model.model.transformer_blocks.0.attn1.to_q.lora_A.weight" and len(names) == 16


@pytest.mark.parametrize("variant", ["t2v", "i2v_posemb", "1.5_temporal_patch"])
def test_training_forward_matches_inference_with_zero_lora(lib, variant):
    """With B = 0 (PEFT's initial state) the differentiable forward equals the inference transformer called without rotary
    embeddings up to the different rounding points of the unfused training path (<= 2e-2 of the max)."""
    from oracle import dit_torch as O
    from videogpa_b200.train_dit import LoRATrainableTransformer
    from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig
    kw = dict(num_attention_heads=4, num_layers=2, text_embed_dim=256, sample_width=24, sample_height=16, sample_frames=9, max_text_seq_length=18)
    cin = 16
    if variant == "i2v_posemb":                               # train/CogVideoX-I2V-5B: in_channels 32, learned positional embedding
        kw.update(in_channels=32, use_learned_positional_embeddings=True)
        cin = 32
    frames = 3
    if variant == "1.5_temporal_patch":                       # train/CogVideoX1.5-5B: patch_size_t 2 (Linear patch embed, even frame count)
        kw.update(patch_size_t=2)
        frames = 4
    sd = {k: v.to(BF).float() for k, v in O.random_state_dict(O.DiTConfig(**kw), seed=8, randomize_norms=True, std=0.05).items()}
    base = CogVideoXTransformer3D(TransformerConfig(**kw), sd, device="cuda")
    pol = LoRATrainableTransformer(base)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(4, frames, cin, 16, 24, generator=g).cuda()
    e = torch.randn(4, 18, 256, generator=g).to(BF).cuda()
    t = torch.tensor([5, 300, 700, 999]).cuda()
    with torch.no_grad():
        a = pol(x, e, t)
        b = base(x, encoder_hidden_states=e, timestep=t).sample
    assert a.shape == b.shape and relmax(a, b) < 2e-2


def test_lora_save_pretrained_round_trip(lib, tmp_path):
    """Train-side `save_pretrained` -> generate-side `--lora_path` merge: the adapter directory written by the trainable wrapper
    is read back by lora.merge_lora and yields W + (alpha / r) B A on the merged transformer (reference 03_train.py:287 ->
    generate/CogVideoX-5B.py:24-31)."""
    from oracle import dit_torch as O
    from videogpa_b200.lora import merge_lora, read_adapter
    from videogpa_b200.train_dit import TARGETS, LoRATrainableTransformer
    from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig
    kw = dict(num_attention_heads=4, num_layers=2, text_embed_dim=256, sample_width=24, sample_height=16, sample_frames=9, max_text_seq_length=18)
    sd = {k: v.to(BF).float() for k, v in O.random_state_dict(O.DiTConfig(**kw), seed=9, std=0.05).items()}
    base = CogVideoXTransformer3D(TransformerConfig(**kw), sd, device="cuda")
    pol = LoRATrainableTransformer(base, r=64, lora_alpha=128.0, seed=1)
    g = torch.Generator(device="cuda").manual_seed(2)
    with torch.no_grad():
        for layer in pol.lora:
            for m in TARGETS:
                layer[m][1].copy_(0.05 * torch.randn(layer[m][1].shape, device="cuda", generator=g))
    out = tmp_path / "final_lora"
    pol.save_pretrained(str(out))
    cfg, pairs = read_adapter(str(out))
    assert cfg["r"] == 64 and cfg["lora_alpha"] == 128.0 and sorted(cfg["target_modules"]) == sorted(TARGETS) and len(pairs) == 8
    fresh = CogVideoXTransformer3D(TransformerConfig(**kw), sd, device="cuda")
    w0 = fresh.attention_weight(1, "to_k").float().clone()
    assert merge_lora(fresh, str(out)) == 8
    want = w0 + pol.merged_delta(1, "to_k")
    assert relmax(fresh.attention_weight(1, "to_k"), want) < 1e-2


def test_train_cli_synthetic_end_to_end(lib, tmp_path, capsys):
    """The 03_train.py mirror on a tiny dataset: metadata JSON + .pt latents on disk -> DPODataset -> split -> 3 optimizer
    steps with accumulation 2, clipping, cosine warm-up -> validation -> final_lora directory that merge_lora reads."""
    import json
    from videogpa_b200.lora import read_adapter
    from videogpa_b200.train import cogvideox_5b as t
    from videogpa_b200.transformer import TransformerConfig
    g = torch.Generator().manual_seed(0)
    base = tmp_path / "data"
    (base / "lat").mkdir(parents=True)
    groups = []
    for gi in range(60):                                            # 60 groups -> 58 train / 2 validation pairs
        vids = []
        for vi, score in enumerate((0.2, 0.9)):
            lp = f"lat/latent_{gi}_{vi}.pt"
            torch.save(torch.randn(16, 3, 16, 24, generator=g), base / lp)
            vids.append({"video_path": f"v{gi}_{vi}.mp4", "consistency_score": score, "motion_norm": 0.5, "latent_path": lp,
                         "condition_path": f"lat/cond_{gi}.pt"})
        torch.save({"encoder_hidden_states": torch.randn(18, 4096, generator=g).to(BF)}, base / f"lat/cond_{gi}.pt")
        groups.append({"group_id": f"g{gi}", "text_prompt": f"prompt {gi}", "videos": vids})
    meta = base / "meta.json"
    meta.write_text(json.dumps({"groups": groups}))
    cfg = dict(t.DEFAULT_CONFIG)
    cfg.update(base_path=str(base), metadata_path=str(meta), output_dir=str(tmp_path / "out"), max_steps=3, warmup_steps=2,
               max_epochs=1, log_every_n_steps=1, learning_rate=1e-4, devices=[0])
    # a 1-block transformer at the 5B width keeps the run short; the latent grid is 3 x 16 x 24
    res = t.main_train(cfg, synthetic_layers=1)
    txt = capsys.readouterr().out
    assert res["steps"] == 3 and len(res["val"]) == 1 and "step 3: train/loss" in txt and "val/loss" in txt
    acfg, pairs = read_adapter(str(tmp_path / "out" / "final_lora"))
    assert acfg["r"] == 64 and len(pairs) == 4                      # one block x to_q / to_k / to_v / to_out.0
    A, Bm = pairs[(0, "to_q")]
    assert float(Bm.abs().max()) > 0                                # B left its zero initialisation


@pytest.mark.parametrize("M,C_,ld", [(306, 256, 256), (612, 64, 768), (37, 130, 130), (128, 64, 64)])
def test_transpose_kernel_bit_exact(lib, M, C_, ld):
    """vgpa_transpose_bf16 (operands of the LoRA weight-gradient GEMMs): exact copy, strided input, zero-padded token columns."""
    from videogpa_b200.train_dit import _t_rows
    g = torch.Generator().manual_seed(M)
    buf = torch.randn(M, ld, generator=g).to(BF).cuda()
    x = buf[:, :C_]
    out = _t_rows(x)
    Mp = (M + 7) // 8 * 8
    assert out.shape == (C_, Mp) and torch.equal(out[:, :M], x.t()) and (Mp == M or float(out[:, M:].abs().max()) == 0.0)


def test_i2v_training_step_with_image_condition(lib):
    """train/CogVideoX-I2V-5B/03_train.py:114-148: the image condition (resize -> VAE encode -> sample x scaling_factor -> first
    latent frame, zero frames after) is channel-concatenated to both noisy latents; the step differentiates to the LoRA factors."""
    from oracle import dit_torch as O
    from oracle import vae_torch as V
    from videogpa_b200.train_dit import LoRATrainableTransformer
    from videogpa_b200.train_step import DPOSharedStep
    from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig
    from videogpa_b200.vae import AutoencoderKLCogVideoXEncoder, VAEDecoderConfig
    kw = dict(num_attention_heads=4, num_layers=2, text_embed_dim=256, sample_width=24, sample_height=16, sample_frames=9, max_text_seq_length=18,
              in_channels=32, use_learned_positional_embeddings=True)
    sd = {k: v.to(BF).float() for k, v in O.random_state_dict(O.DiTConfig(**kw), seed=10, randomize_norms=True, std=0.05).items()}
    base = CogVideoXTransformer3D(TransformerConfig(**kw), sd, device="cuda")
    vcfg = VAEDecoderConfig(block_out_channels=(64, 64, 64, 128), sample_height=256, sample_width=384)
    enc = AutoencoderKLCogVideoXEncoder(_bf16(V.random_encoder_state_dict(V.VAEConfig(block_out_channels=(64, 64, 64, 128)), seed=30)), vcfg, device="cuda")
    pol = LoRATrainableTransformer(base, gradient_checkpointing="mlp")
    with torch.no_grad():
        for layer in pol.lora:
            for m in layer:
                layer[m][1].normal_(0, 0.02)
    step = DPOSharedStep(base, None, beta=20.0, trainable=pol, vae_encoder=enc)
    g = torch.Generator().manual_seed(5)
    batch = {"x_win": torch.randn(2, 16, 3, 16, 24, generator=g), "x_lose": torch.randn(2, 16, 3, 16, 24, generator=g),
             "prompt_emb": torch.randn(2, 18, 256, generator=g).to(BF), "image_emb": torch.rand(2, 3, 100, 150, generator=g) * 2 - 1}
    x_win = batch["x_win"].cuda().permute(0, 2, 1, 3, 4).float()
    gen = torch.Generator(device="cuda").manual_seed(3)
    cond = step._image_condition(batch, x_win, gen)
    assert cond.shape == x_win.shape and float(cond[:, 1:].abs().max()) == 0.0 and float(cond[:, 0].abs().max()) > 0
    img = torch.nn.functional.interpolate(batch["image_emb"].cuda().float(), size=(128, 192))
    want = enc.encode(img.unsqueeze(2).to(BF)).latent_dist.sample(generator=torch.Generator(device="cuda").manual_seed(3)) * 0.7
    assert torch.equal(cond[:, 0], want[:, :, 0].float())                      # [B, C, 1, h, w] -> first latent frame
    del batch["image_emb"]
    assert float(step._image_condition(batch, x_win).abs().max()) == 0.0       # no image: zeros_like
    batch["image_emb"] = torch.rand(2, 3, 100, 150, generator=g) * 2 - 1
    loss = step.training_step(batch, generator=torch.Generator(device="cuda").manual_seed(9))
    loss.backward()
    assert torch.isfinite(loss) and all(p.grad is not None and torch.isfinite(p.grad).all() for p in pol.parameters())


def _bf16(sd):
    return {k: v.to(BF).float() for k, v in sd.items()}


def test_cogvideox1_5_training_step_trims_and_matches_oracle(lib):
    """train/CogVideoX1.5-5B/03_train.py:134-157 on a temporal-patch model: odd latent sizes are trimmed to even F / H / W, the
    loss of training_step equals the oracle's (fp32 CPU forward of the same 2-block model, LoRA at its PEFT initial state so
    policy == reference and the loss is log 2 exactly), and backward reaches the LoRA factors with finite, non-zero dA."""
    from oracle import dit_torch as O
    from videogpa_b200.train_dit import LoRATrainableTransformer
    from videogpa_b200.train_step import DPOSharedStep
    from videogpa_b200.transformer import CogVideoXTransformer3D, TransformerConfig
    kw = dict(num_attention_heads=4, num_layers=2, text_embed_dim=256, sample_width=24, sample_height=16, sample_frames=9, max_text_seq_length=18,
              patch_size_t=2)
    ocfg = O.DiTConfig(**kw)
    sd = {k: v.to(BF).float() for k, v in O.random_state_dict(ocfg, seed=9, randomize_norms=True, std=0.05).items()}
    base = CogVideoXTransformer3D(TransformerConfig(**kw), sd, device="cuda")
    pol = LoRATrainableTransformer(base, r=64, lora_alpha=128.0, seed=3)
    g = torch.Generator().manual_seed(5)
    Bsz, C, Fr, H, W, St = 1, 16, 5, 17, 25, 18                   # odd everywhere: trimmed to 4 x 16 x 24
    batch = {"x_win": torch.randn(Bsz, C, Fr, H, W, generator=g), "x_lose": torch.randn(Bsz, C, Fr, H, W, generator=g),
             "prompt_emb": torch.randn(Bsz, St, 256, generator=g).to(BF)}
    t = torch.tensor([640])
    noise = torch.randn(Bsz, Fr, C, H, W, generator=g)
    step = DPOSharedStep(base, None, beta=1.0, trainable=pol)
    loss = step.training_step(batch, timesteps=t.cuda(), noise=noise.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - math.log(2.0)) < 2e-3                # B = 0: policy == reference up to the training path's rounding points
    # the prediction itself against the oracle on the trimmed input
    ac = O.cogvideox_alphas_cumprod()
    xw = batch["x_win"].permute(0, 2, 1, 3, 4)[:, :4, :, :16, :24]
    nz = noise[:, :4, :, :16, :24]
    want = O.transformer_forward(sd, ocfg, O.add_noise(ac, xw, nz, t.numpy()), batch["prompt_emb"].float(), t, None)
    with torch.no_grad():
        got = pol(O.add_noise(ac, xw, nz, t.numpy()).cuda(), batch["prompt_emb"].cuda(), t.cuda())
    assert got.shape == want.shape == (1, 4, 16, 16, 24) and relmax(got.cpu(), want) < 3e-2
    grads = [p.grad for n, p in pol.named_parameters() if "lora_B" in n]
    assert all(gr is not None and torch.isfinite(gr).all() for gr in grads) and any(float(gr.abs().max()) > 0 for gr in grads)


def test_train_cli_1_5_synthetic(lib, tmp_path, capsys):
    """`python -m videogpa_b200.train.cogvideox1_5_5b --synthetic 1`: the 1.5 defaults (max_steps 1500, no VAE switches) and one
    optimizer step on a 1-block temporal-patch model fed by a dataset with odd latent sizes."""
    import json
    import yaml
    from videogpa_b200.train import cogvideox1_5_5b as cli
    assert cli.DEFAULT_CONFIG["max_steps"] == 1500 and cli.DEFAULT_CONFIG["model_path"] == "THUDM/CogVideoX1.5-5B"
    assert "enable_tiling" not in cli.DEFAULT_CONFIG
    root = tmp_path / "data"
    (root / "lat").mkdir(parents=True)
    g = torch.Generator().manual_seed(0)
    groups = []
    for gi in range(3):
        vids = []
        for vi, score in enumerate((0.1, 0.9)):
            lp = f"lat/g{gi}_v{vi}.pt"
            torch.save(torch.randn(16, 3, 9, 13, generator=g), str(root / lp))
            cp = f"lat/g{gi}_v{vi}_cond.pt"
            torch.save({"encoder_hidden_states": torch.randn(18, 4096, generator=g).to(BF)}, str(root / cp))
            vids.append({"video_path": f"v{gi}_{vi}.mp4", "consistency_score": score, "motion_norm": 1.0, "latent_path": lp, "condition_path": cp})
        groups.append({"group_id": f"g{gi}", "text_prompt": "p", "videos": vids})
    (root / "meta.json").write_text(json.dumps({"groups": groups}))
    cfgf = tmp_path / "c.yaml"
    cfgf.write_text(yaml.safe_dump({"training": {"metadata_path": str(root / "meta.json"), "output_dir": str(tmp_path / "out"), "max_steps": 1,
                                                 "accumulate_grad_batches": 1, "warmup_steps": 1, "devices": [0], "log_every_n_steps": 1}}))
    res = cli.main(["--config", str(cfgf), "--base_path", str(root), "--synthetic", "1"])
    assert res["steps"] == 1 and math.isfinite(res["last_loss"])
    assert (tmp_path / "out" / "final_lora" / "adapter_model.safetensors").exists()
