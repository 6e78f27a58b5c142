"""GPU parity of the T5 prompt encoder against the REAL third-party implementation the reference calls
— transformers' T5EncoderModel (installed in this image) — on seeded random weights of a small v1.1-style config. The
transformers model runs on the CPU in fp32 with bf16-rounded weights; tolerance 3e-2 of the max for the whole encoder
(bf16 activations in the CUDA path), 1e-2 for the single attention op against an fp32 torch restatement."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def relmax(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


def _hf_model(seed=0, **kw):
    from transformers import T5Config as HFConfig, T5EncoderModel as HFEncoder
    cfg = HFConfig(vocab_size=100, d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4, feed_forward_proj="gated-gelu",
                   relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.0, **kw)
    torch.manual_seed(seed)
    m = HFEncoder(cfg).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "layer_norm" in n:
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            elif "relative_attention_bias" in n:
                p.copy_(0.5 * torch.randn_like(p))
            elif "SelfAttention.q" in n:
                p.copy_(torch.randn_like(p) * (256 * 64) ** -0.5 * 4)     # scores of O(1): T5 has no 1/sqrt(d) scaling
            elif "shared" in n or "embed_tokens" in n:
                p.copy_(torch.randn_like(p))
            else:
                p.copy_(torch.randn_like(p) * p.shape[1] ** -0.5)
            p.copy_(p.to(BF).float())
    return cfg, m


def _mine(cfg, m):
    from videogpa_b200.t5 import T5Config, T5EncoderModel
    c = T5Config(vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff, num_layers=cfg.num_layers,
                 num_heads=cfg.num_heads, relative_attention_num_buckets=cfg.relative_attention_num_buckets,
                 relative_attention_max_distance=cfg.relative_attention_max_distance, layer_norm_epsilon=cfg.layer_norm_epsilon)
    return T5EncoderModel(c, m.state_dict(), device="cuda")


@pytest.mark.parametrize("B,S", [(1, 226), (2, 37)])
def test_t5_encoder_vs_transformers(lib, B, S):
    cfg, m = _hf_model()
    enc = _mine(cfg, m)
    ids = torch.randint(0, 100, (B, S), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = m(ids)[0]
    out = enc(ids.cuda())
    got = out[0].cpu().float()
    assert got.shape == ref.shape == (B, S, 256) and out.last_hidden_state is out[0]
    assert relmax(got, ref) < 3e-2, relmax(got, ref)
    assert (got - ref).abs().mean().item() < 5e-3 * ref.abs().max().item()


def test_umt5_encoder_vs_transformers(lib):
    """umT5 (the text encoder of Wan2.2: one relative-position table per block) against transformers' UMT5EncoderModel, without a mask
    and with Wan's padded-prompt-plus-mask call: the rows of the real tokens must match, padded rows come back as zeros."""
    from transformers import UMT5Config, UMT5EncoderModel
    from videogpa_b200.t5 import T5Config, T5EncoderModel
    cfg = UMT5Config(vocab_size=100, d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4, feed_forward_proj="gated-gelu", dropout_rate=0.0)
    torch.manual_seed(5)
    m = UMT5EncoderModel(cfg).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "layer_norm" in n:
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            elif "relative_attention_bias" in n:
                p.copy_(0.5 * torch.randn_like(p))
            elif "SelfAttention.q" in n:
                p.copy_(torch.randn_like(p) * (256 * 64) ** -0.5 * 4)
            elif "shared" in n or "embed_tokens" in n:
                p.copy_(torch.randn_like(p))
            else:
                p.copy_(torch.randn_like(p) * p.shape[1] ** -0.5)
            p.copy_(p.to(BF).float())
    sd = m.state_dict()
    tables = [sd[f"encoder.block.{i}.layer.0.SelfAttention.relative_attention_bias.weight"] for i in range(2)]
    assert not torch.equal(tables[0], tables[1])                                      # the per-block tables really differ
    c = T5Config(vocab_size=100, d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4, per_layer_relative_bias=True)
    enc = T5EncoderModel(c, sd, device="cuda")
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(1, 100, (2, 40), generator=g)
    with torch.no_grad():
        ref = m(ids)[0]
    got = enc(ids.cuda())[0].cpu().float()
    err = (got - ref).abs().mean().item()
    assert relmax(got, ref) < 3e-2 and err < 5e-3 * ref.abs().max().item(), (relmax(got, ref), err)
    # a shared-table (plain T5) reading of the same weights must NOT match: the per-block tables matter
    wrong = T5EncoderModel(T5Config(vocab_size=100, d_model=256, d_kv=64, d_ff=512, num_layers=2, num_heads=4), sd, device="cuda")
    err_wrong = (wrong(ids.cuda())[0].cpu().float() - ref).abs().mean().item()
    assert err_wrong > 5 * err, (err_wrong, err)
    # Wan's call: padded to a fixed length with a mask; lengths 40 and 23
    mask = torch.ones(2, 40, dtype=torch.long); mask[1, 23:] = 0
    ids_p = ids.clone(); ids_p[1, 23:] = 0
    with torch.no_grad():
        ref_m = m(ids_p, attention_mask=mask)[0]
    got_m = enc(ids_p.cuda(), attention_mask=mask.cuda())[0].cpu().float()
    assert relmax(got_m[0], ref_m[0]) < 3e-2 and relmax(got_m[1, :23], ref_m[1, :23]) < 3e-2
    assert float(got_m[1, 23:].abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        bad = mask.clone(); bad[0, 3] = 0                                              # a hole in the mask is not right padding
        enc(ids_p.cuda(), attention_mask=bad.cuda())
    assert T5Config.umt5_xxl().vocab_size == 256384 and T5Config.umt5_xxl().per_layer_relative_bias


@pytest.mark.parametrize("B,H,S", [(2, 3, 75), (1, 2, 1), (1, 1, 512), (1, 4, 226), (3, 1, 33)])
def test_t5_attention_op_vs_torch(lib, B, H, S):
    """vgpa_t5_attention_bf16 alone: no 1/sqrt(d) scaling, additive bias, bf16 roundings of scores and probabilities."""
    from videogpa_b200 import _lib
    L = _lib.load()
    g = torch.Generator().manual_seed(2)
    qkv = (torch.randn(B * S, 3 * H * 64, generator=g) * 0.35).to(BF)
    bias = (torch.randn(H, S, S, generator=g) * 0.5).to(BF)
    q, k, v = [t.float().view(B, S, H, 64).permute(0, 2, 1, 3) for t in qkv.split(H * 64, dim=1)]
    sc = (q @ k.transpose(-1, -2)).to(BF).float()
    sc = (sc + bias.float()[None]).to(BF).float()
    p = torch.softmax(sc, dim=-1).to(BF).float()
    ref = (p @ v).permute(0, 2, 1, 3).reshape(B * S, H * 64)
    dq = qkv.cuda()
    out = torch.empty(B * S, H * 64, dtype=BF, device="cuda")
    esz = 2
    _lib.check(L.vgpa_t5_attention_bf16(dq.data_ptr(), dq.data_ptr() + H * 64 * esz, dq.data_ptr() + 2 * H * 64 * esz,
                                        bias.cuda().data_ptr(), out.data_ptr(), B, H, S, dq.stride(0), out.stride(0),
                                        _lib.current_stream()), "vgpa_t5_attention_bf16")
    assert relmax(out.cpu(), ref) < 1e-2, relmax(out.cpu(), ref)


def test_gated_mul_bit_exact(lib):
    from videogpa_b200 import _lib
    L = _lib.load()
    g = torch.Generator().manual_seed(3)
    a = torch.randn(37, 512, generator=g).to(BF).cuda()
    b = torch.randn(37, 512, generator=g).to(BF).cuda()
    out = torch.empty_like(a)
    _lib.check(L.vgpa_gated_mul_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), 37, 512, 512, 512, 512, _lib.current_stream()), "gated_mul")
    assert torch.equal(out, a * b)                         # one bf16 rounding of the exact product, as eager torch


def test_t5_rejects_bad_input(lib):
    cfg, m = _hf_model()
    enc = _mine(cfg, m)
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 8, dtype=torch.long))                                   # CPU ids
    with pytest.raises(RuntimeError):
        enc(torch.full((1, 8), 100, dtype=torch.long, device="cuda"))              # id out of range
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 8, dtype=torch.long, device="cuda"), attention_mask=torch.tensor([[1, 1, 1, 1, 0, 0, 0, 0]], device="cuda"))
    sd = m.state_dict()
    del sd["encoder.final_layer_norm.weight"]
    from videogpa_b200.t5 import T5Config, T5EncoderModel
    with pytest.raises(RuntimeError):
        T5EncoderModel(T5Config(vocab_size=100, d_model=256, d_ff=512, num_layers=2, num_heads=4), sd, device="cuda")


def test_encode_text_condition_mirror_and_graph_replay(lib):
    """02_encode.py:69-93 tensor half; a second prompt of the same length replays the captured graph and must equal eager."""
    from videogpa_b200.encode import encode_text_condition
    cfg, m = _hf_model(seed=5)
    enc = _mine(cfg, m)
    g = torch.Generator().manual_seed(9)
    ids1 = torch.randint(0, 100, (1, 226), generator=g)
    ids2 = torch.randint(0, 100, (1, 226), generator=g)
    c1 = encode_text_condition(enc, ids1)
    c2 = encode_text_condition(enc, ids2)                    # graph replay with new ids
    assert set(c1) == {"encoder_hidden_states"} and c1["encoder_hidden_states"].shape == (226, 256) and c1["encoder_hidden_states"].device.type == "cpu"
    enc.use_cuda_graph = False
    e2 = enc(ids2.cuda())[0][0].cpu()
    assert torch.equal(c2["encoder_hidden_states"], e2) and not torch.equal(c1["encoder_hidden_states"], e2)
    with torch.no_grad():
        ref = m(ids2)[0][0]
    assert relmax(e2, ref) < 3e-2
