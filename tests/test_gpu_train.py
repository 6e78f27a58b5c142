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


@pytest.mark.parametrize("B,H,Sq,Skv", [(1, 2, 128, 128), (2, 3, 300, 300), (1, 2, 200, 450)])
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
