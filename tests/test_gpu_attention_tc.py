"""GPU: attention on tcgen05 / TMEM (csrc/attention_tc.cu) against a plain PyTorch fp32 statement of
Attention.forward (/root/reference/models/act.py:57-66) and its autograd backward: packed short sequences (several
sequences per 128-row tile, block-diagonal scores), ragged last tiles, long sequences (flash loop over K/V tiles)."""
import pytest
import torch

from act_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def force_tcgen05_attention():
    """act_set_option(ACT_OPT_ATTN_TC, 2): the tcgen05 kernels for every length they implement (the default dispatch sends
    the latency-bound short sequences to the warp-MMA kernels, which tests/test_gpu_layers.py covers)."""
    from act_b200 import _lib
    assert _lib.lib().act_set_option(2, 2) == 0
    yield
    assert _lib.lib().act_set_option(2, 1) == 0


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _ref(qkv, B, T, H, scale):
    r = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)                      # 3 B H T 64
    s = (r[0] @ r[1].transpose(-2, -1)) * scale
    o = (s.softmax(-1) @ r[2]).transpose(1, 2).reshape(B * T, H * 64)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,T,H", [(128, 27, 6), (7, 27, 6), (128, 64, 6), (5, 65, 6), (3, 128, 12), (9, 14, 6), (1, 1, 6),
                                   (16, 206, 6), (4, 512, 6), (3, 300, 2), (2, 129, 6)])
def test_attention_tc_forward(B, T, H):
    torch.manual_seed(B * 1000 + T)
    qkv = (torch.randn(B * T, 3 * H * 64, device="cuda") * 0.8).bfloat16()
    o, lse = ops.attention_fwd(qkv, B, T, H, 0.125)
    want, wl = _ref(qkv, B, T, H, 0.125)
    assert torch.isfinite(o.float()).all()
    assert rel(o, want) < 5e-3, rel(o, want)
    torch.testing.assert_close(lse, wl, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("B,T,H", [(128, 27, 6), (7, 27, 6), (128, 64, 6), (5, 65, 6), (3, 128, 6), (9, 14, 6), (2, 1, 6),
                                   (16, 206, 6), (4, 512, 6), (3, 300, 2), (2, 129, 6)])
def test_attention_tc_backward(B, T, H):
    torch.manual_seed(B * 77 + T)
    qkv = (torch.randn(B * T, 3 * H * 64, device="cuda") * 0.8).bfloat16()
    do = (torch.randn(B * T, H * 64, device="cuda") * 0.5).bfloat16()
    o, lse = ops.attention_fwd(qkv, B, T, H, 0.125)
    dqkv = ops.attention_bwd(qkv, o, do, lse, B, T, H, 0.125)
    x = qkv.float().requires_grad_(True)
    want, _ = _ref(x, B, T, H, 0.125)
    want.backward(do.float())
    assert torch.isfinite(dqkv.float()).all()
    g = x.grad.view(B * T, 3, H * 64)
    d = dqkv.float().view(B * T, 3, H * 64)
    for i, name in enumerate("qkv"):
        if g[:, i].norm().item() < 1e-6:                     # T = 1: softmax of one score, dq = dk = 0 exactly
            assert d[:, i].abs().max().item() < 1e-3, name
            continue
        assert rel(d[:, i], g[:, i]) < 1.5e-2, (name, rel(d[:, i], g[:, i]))
