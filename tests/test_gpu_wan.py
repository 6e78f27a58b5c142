"""GPU parity of the Wan2.2 DiT path against the torch oracle (oracle/wan_torch.py, parity
unpinned against the un-vendored Wan2.2 repo — see its header). Tolerances: single ops <= 1e-2 of the max against fp32
torch on bf16-rounded inputs; the 2-block model forward <= 5e-2 of the max (bf16 residual stream vs fp32 oracle)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import wan_torch as O

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def relmax(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("B,H,Sq,Skv", [(1, 1, 128, 128), (1, 2, 300, 300), (2, 3, 1000, 1000), (1, 2, 700, 64), (1, 2, 333, 512)])
def test_attention_head_dim_128_vs_sdpa(lib, B, H, Sq, Skv):
    from videogpa_b200 import dense
    g = torch.Generator().manual_seed(Sq + Skv)
    D = H * 128
    q = torch.randn(B, Sq, D, generator=g).to(BF).cuda()
    kv = torch.randn(B, Skv, 2 * D, generator=g).to(BF).cuda()
    out = dense.attention(q, kv[..., :D], kv[..., D:], H, head_dim=128)
    sp = lambda t, n: t.reshape(B, n, H, 128).transpose(1, 2).float()
    ref = F.scaled_dot_product_attention(sp(q, Sq), sp(kv[..., :D], Skv), sp(kv[..., D:], Skv)).transpose(1, 2).reshape(B, Sq, D)
    assert torch.isfinite(out.float()).all() and relmax(out, ref) < 1e-2


def test_rmsnorm_rope_vs_oracle(lib):
    from videogpa_b200.wan import WanConfig, _rmsnorm_rope, rope_tables
    cfg = WanConfig(dim=512, num_heads=4)
    ocfg = O.WanConfig(dim=512, num_heads=4)
    g = torch.Generator().manual_seed(2)
    f, h, w = 3, 4, 5
    S = f * h * w
    buf = torch.randn(S, 3 * 512, generator=g).to(BF)
    wt = 1 + 0.1 * torch.randn(512, generator=g)
    cos, sin = rope_tables(cfg, f, h, w, device="cuda")
    oc, os_ = O.rope_tables(ocfg, f, h, w)
    assert torch.equal(cos.cpu(), oc) and torch.equal(sin.cpu(), os_)
    x = buf.cuda()
    _rmsnorm_rope(x[:, 512:1024], wt.cuda(), 1e-6, (cos, sin), 128)          # the k slice of a fused projection, in place
    ref = O.rope_apply(O.rms_norm(buf[:, 512:1024], wt, 1e-6).view(S, 4, 128), oc, os_).reshape(S, 512)
    assert relmax(x[:, 512:1024].cpu(), ref) < 1e-2
    assert torch.equal(x[:, :512].cpu(), buf[:, :512]) and torch.equal(x[:, 1024:].cpu(), buf[:, 1024:])
    y = buf[:, :512].contiguous().cuda()
    _rmsnorm_rope(y, wt.cuda(), 1e-6)                                         # no rope: cross-attention q / k
    assert relmax(y.cpu(), O.rms_norm(buf[:, :512], wt, 1e-6)) < 1e-2


def _small():
    kw = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128, text_len=64)
    return kw


def test_wan_forward_vs_oracle(lib):
    from videogpa_b200.wan import WanConfig, WanTransformer3D
    kw = _small()
    ocfg = O.WanConfig(**kw)
    sd = {k: v.to(BF).float() for k, v in O.random_state_dict(ocfg, seed=3, std=0.05).items()}
    model = WanTransformer3D(WanConfig(**kw), sd, device="cuda")
    g = torch.Generator().manual_seed(4)
    lat = torch.randn(48, 3, 8, 12, generator=g).to(BF)
    ctx = torch.randn(40, 128, generator=g).to(BF)
    S, hw = 3 * 4 * 6, 4 * 6
    t = torch.full((S,), 850.0); t[:hw] = 0.0                                 # TI2V: first latent frame carries t = 0
    out = model([lat.cuda()], t[None], [ctx.cuda()], seq_len=S)[0].cpu().float()
    ref = O.model_forward(sd, ocfg, lat.float(), t, ctx.float())
    assert out.shape == ref.shape == (48, 3, 8, 12)
    assert relmax(out, ref) < 5e-2, relmax(out, ref)
    # scalar timestep (T2V-style) and the tensor call form
    out2 = model(lat.cuda()[None], torch.tensor([500.0]), ctx.cuda()[None])[0].cpu().float()
    ref2 = O.model_forward(sd, ocfg, lat.float(), 500.0, ctx.float())
    assert relmax(out2, ref2) < 5e-2
    with pytest.raises(RuntimeError):
        bad = t.clone(); bad[hw + 1] = 3.0
        model([lat.cuda()], bad[None], [ctx.cuda()])


def test_wan_denoise_step_vs_oracle(lib):
    from videogpa_b200.wan import WanConfig, WanDenoiseStep, WanTransformer3D, flow_sigmas
    kw = _small()
    ocfg = O.WanConfig(**kw)
    sd = {k: v.to(BF).float() for k, v in O.random_state_dict(ocfg, seed=5, std=0.05).items()}
    model = WanTransformer3D(WanConfig(**kw), sd, device="cuda")
    sig = flow_sigmas(50, shift=5.0)
    osig = O.flow_sigmas(50, shift=5.0)
    assert max(abs(a - float(b)) for a, b in zip(sig, osig)) < 1e-12 and sig[0] == 1.0 and sig[-1] == 0.0
    g = torch.Generator().manual_seed(6)
    lat = torch.randn(48, 2, 8, 8, generator=g).to(BF)
    ctx, ctx0 = torch.randn(30, 128, generator=g).to(BF), torch.zeros(1, 128).to(BF)
    step = WanDenoiseStep(model, guide_scale=5.0)
    tt = sig[3] * 1000
    nxt = step(lat.cuda(), torch.tensor([tt]), sig[3], sig[4], ctx.cuda(), ctx0.cuda(), first_frame=lat[:, :1].cuda()).cpu().float()
    c = O.model_forward(sd, ocfg, lat.float(), tt, ctx.float())
    u = O.model_forward(sd, ocfg, lat.float(), tt, ctx0.float())
    ref = O.euler_flow_step(lat.float(), u + 5.0 * (c - u), sig[3], sig[4])
    ref[:, :1] = lat[:, :1].float()
    assert relmax(nxt, ref) < 3e-2
    assert torch.equal(nxt[:, :1], lat[:, :1].float())


def test_wan_lora_merge_vs_oracle(lib, tmp_path):
    """`--lora_path` for Wan: adapters target q, k, v, o of self_attn and cross_attn (train/Wan2.2-TI2V-5B/03_train.py:82) and the
    generate script scales alpha/r by --lora_weight 0.2 (generate/Wan2.2-TI2V-5B.py:66-70)."""
    import json
    from safetensors.torch import save_file
    from oracle import dit_torch as OD
    from videogpa_b200.lora import merge_lora
    from videogpa_b200.wan import WanConfig, WanTransformer3D
    kw = _small()
    ocfg = O.WanConfig(**kw)
    sd = {k: v.to(BF) for k, v in O.random_state_dict(ocfg, seed=8).items()}
    model = WanTransformer3D(WanConfig(**kw), sd, device="cuda")
    D, r = ocfg.dim, 64
    g = torch.Generator().manual_seed(10)
    tensors, expect = {}, {}
    for layer in range(2):
        for att in ("self_attn", "cross_attn"):
            for m in "qkvo":
                A = torch.randn(r, D, generator=g) * 0.05
                Bm = torch.randn(D, r, generator=g) * 0.05
                base = f"base_model.model.blocks.{layer}.{att}.{m}"
                tensors[base + ".lora_A.weight"], tensors[base + ".lora_B.weight"] = A, Bm
                expect[(layer, f"{att}.{m}")] = OD.lora_merge(sd[f"blocks.{layer}.{att}.{m}.weight"], A, Bm, (128.0 / 64.0) * 0.2)
    save_file(tensors, str(tmp_path / "adapter_model.safetensors"))
    (tmp_path / "adapter_config.json").write_text(json.dumps(dict(peft_type="LORA", r=64, lora_alpha=128.0, target_modules=["q", "k", "v", "o"],
                                                                use_dora=False, use_rslora=False, fan_in_fan_out=False, bias="none")))
    assert merge_lora(model, str(tmp_path), weight=0.2) == 16
    for (layer, mod), ref in expect.items():
        got = model.attention_weight(layer, mod).cpu()
        diff = (got.float() - ref.float()).abs()
        assert (got != ref).float().mean() < 0.01 and diff.max() <= 2.0 ** -7 * ref.float().abs().max()


def test_wan_generate_cli_synthetic(lib, tmp_path, capsys):
    """generate/Wan2.2-TI2V-5B.py surface on the GPU with a 1-block random DiT: image resolution through --base_dir, missing
    image skipped, skip-if-exists on the second run, final latent with the first frame clamped to the image latent."""
    import json
    from videogpa_b200.generate import wan2_2_ti2v_5b as g
    (tmp_path / "imgs").mkdir()
    (tmp_path / "imgs" / "a.png").write_bytes(b"bytes that seed the synthetic first-frame latent")
    pj = tmp_path / "p.json"
    pj.write_text(json.dumps({"s1": {"text_prompt": "a boat", "image_prompt": "a.png"}, "s2": {"text_prompt": "a car", "image_prompt": "missing.png"},
                              "s3": {"text_prompt": "no image"}}))
    out = tmp_path / "out"
    argv = ["--model_path", "unused", "--prompt_json", str(pj), "--output_dir", str(out), "--base_dir", str(tmp_path / "imgs"),
            "--synthetic", "1", "--sampling_steps", "3", "--frame_num", "9", "--height", "128", "--width", "192"]
    g.main(argv)
    txt = capsys.readouterr().out
    assert "Failed" not in txt and "Image not found" in txt, txt
    lat = torch.load(str(out / "s1" / "seed_42.latents.pt"))
    assert lat.shape == (48, 3, 8, 12) and lat.dtype == torch.bfloat16 and torch.isfinite(lat.float()).all()
    first = g._seeded(str(tmp_path / "imgs" / "a.png"), (48, 1, 8, 12), "img").to(torch.bfloat16)
    assert torch.equal(lat[:, :1], first)                       # TI2V: the image frame is kept
    assert not (out / "s2").exists() and not (out / "s3").exists()
    g.main(argv)
    assert "Skip existing: s1" in capsys.readouterr().out


def test_fp32_residual_kernels_vs_torch(lib):
    """The two kernels behind the fp32 residual stream of the Wan2.2 forward (torch.autocast semantics): LayerNorm + modulation
    on a float32 row, one bf16 rounding; gated residual update out_f32 += gate * bf16(acc + bias)."""
    from videogpa_b200 import dense
    g = torch.Generator(device="cuda").manual_seed(7)
    S, D, hw = 300, 3072, 40
    x = torch.randn(S, D, generator=g, device="cuda") * 3 + 0.5
    shift = [(torch.randn(1, D, generator=g, device="cuda") * 0.2).to(BF) for _ in range(2)]
    scale = [(torch.randn(1, D, generator=g, device="cuda") * 0.2).to(BF) for _ in range(2)]
    n = dense.layernorm_modulate(x, None, None, eps=1e-6, rows_per_sample=S, text_rows=hw, shift_txt=shift[0], scale_txt=scale[0],
                                 shift_vid=shift[1], scale_vid=scale[1], mod_stride_b=0)
    ln = torch.nn.functional.layer_norm(x, (D,), eps=1e-6)
    is_first = (torch.arange(S, device="cuda") < hw)[:, None]
    ref = ln * (1 + torch.where(is_first, scale[0].float(), scale[1].float())) + torch.where(is_first, shift[0].float(), shift[1].float())
    assert n.dtype == BF and relmax(n, ref) < 5e-3                                    # one bf16 rounding
    w, b = (1 + 0.1 * torch.randn(D, generator=g, device="cuda")).to(BF), (0.1 * torch.randn(D, generator=g, device="cuda")).to(BF)
    n2 = dense.layernorm_modulate(x, w, b, eps=1e-6)
    assert relmax(n2, torch.nn.functional.layer_norm(x, (D,), w.float(), b.float(), 1e-6)) < 5e-3
    # gated residual on the fp32 stream
    a = (torch.randn(S, 512, generator=g, device="cuda") * 0.5).to(BF)
    wt = (torch.randn(D, 512, generator=g, device="cuda") * 0.05).to(BF)
    bias = (torch.randn(D, generator=g, device="cuda") * 0.1).to(BF)
    gate = [(torch.randn(1, D, generator=g, device="cuda")).to(BF) for _ in range(2)]
    out = x.clone()
    dense.linear(a, wt, bias, out=out, epilogue=dense.EPI_GATE_RES_F32, rows_per_sample=S, text_rows=hw, gate_txt=gate[0], gate_vid=gate[1],
                 gate_stride_b=0)
    y = (a.float() @ wt.float().t() + bias.float()).to(BF).float()
    want = x + torch.where(is_first, gate[0].float(), gate[1].float()) * y
    assert out.dtype == torch.float32 and (out - want).abs().max().item() < 2e-2 * y.abs().max().item()   # y differs by <= 1 bf16 ulp
    out2 = x.clone()
    dense.linear(a, wt, bias, out=out2, epilogue=dense.EPI_GATE_RES_F32)                                   # no gate = 1
    assert (out2 - (x + y)).abs().max().item() < 2e-2 * y.abs().max().item()
    with pytest.raises(RuntimeError):
        dense.linear(a, wt, bias, out=x.to(BF), epilogue=dense.EPI_GATE_RES_F32)                           # the stream must be float32
