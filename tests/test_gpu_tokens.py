"""GPU: the token-plumbing kernels of the student path (csrc/tokens.cu) against plain PyTorch fp32 statements of the
reference's own operations (models/act.py:173-177, 276-290, 1219-1229; timm DropPath)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from act_b200 import layers, ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("R,C", [(3328, 384), (8192, 768), (5, 384)])
def test_pos_mlp_forward_backward_vs_torch(R, C):
    torch.manual_seed(R)
    seq = torch.nn.Sequential(torch.nn.Linear(3, 128), torch.nn.GELU(), torch.nn.Linear(128, C)).cuda()
    x = torch.randn(R, 3, device="cuda")
    want = seq(x)
    g = torch.randn_like(want)
    want.backward(g)
    ref = [p.grad.clone() for p in seq.parameters()]
    for p in seq.parameters():
        p.grad = None
    got = layers.pos_mlp(seq, x)
    assert rel(got, want) < 6e-3                                   # bf16 operands of the 128 -> C GEMM
    got.backward(g)
    for p, r in zip(seq.parameters(), ref):
        assert rel(p.grad, r) < 2e-2, p.shape
    # first layer alone, fp32 output: the K = 3 kernel is fp32-exact up to summation order
    a = ops.pos_mlp1_fwd(x, seq[0].weight, seq[0].bias, out_dtype=torch.float32)
    torch.testing.assert_close(a, F.gelu(F.linear(x, seq[0].weight, seq[0].bias)), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,G,ratio", [(128, 64, 0.6), (16, 512, 0.6), (3, 100, 0.8), (2, 64, 0.0)])
def test_mask_order_and_permute_groups(B, G, ratio):
    rng = np.random.default_rng(G)
    nm = int(ratio * G)
    mask = np.zeros((B, G), bool)
    for b in range(B):
        mask[b, rng.permutation(G)[:nm]] = True
    m = torch.from_numpy(mask).cuda()
    order = ops.mask_order(m)
    want = torch.argsort(m.to(torch.uint8), dim=1, stable=True)
    assert torch.equal(order, want)
    k, n_vis = 32, G - nm
    nb = torch.randn(B, G, k, 3, device="cuda")
    center = torch.randn(B, G, 3, device="cuda")
    nb_perm, cs, vc = ops.permute_groups(nb, center, order, n_vis)
    nbs = torch.gather(nb, 1, want[:, :, None, None].expand(-1, -1, k, 3))
    assert torch.equal(nb_perm, torch.cat([nbs[:, :n_vis].reshape(-1, k, 3), nbs[:, n_vis:].reshape(-1, k, 3)], 0))
    css = torch.gather(center, 1, want[..., None].expand(-1, -1, 3))
    assert torch.equal(cs, css) and torch.equal(vc, css[:, :n_vis].reshape(-1, 3))
    _, cs2, vc2 = ops.permute_groups(None, center, order, n_vis, want_nb=False)
    assert torch.equal(cs2, css) and torch.equal(vc2, vc)


def test_assemble_rows_forward_backward():
    torch.manual_seed(0)
    B, n, C = 7, 26, 384
    src = torch.randn(B, n, C, device="cuda", requires_grad=True)
    cls = torch.randn(1, 1, C, device="cuda", requires_grad=True)
    out = layers.assemble_rows(src, cls, B, n, n + 1, True)
    want = torch.cat([cls.expand(B, -1, -1), src], 1)
    assert torch.equal(out, want)
    g = torch.randn_like(out)
    out.backward(g)
    assert torch.equal(src.grad, g[:, 1:]) and rel(cls.grad, g[:, :1].sum(0, keepdim=True)) < 1e-6
    # mask tokens behind the encoder output's visible rows, reading past its cls row
    enc = torch.randn(B, n + 1, C, device="cuda", requires_grad=True)
    tok = torch.randn(1, 1, C, device="cuda", requires_grad=True)
    T = 64
    out = layers.assemble_rows(enc, tok, B, n, T, False, src_off=1)
    assert torch.equal(out, torch.cat([enc[:, 1:], tok.expand(B, T - n, -1)], 1))
    g = torch.randn_like(out)
    out.backward(g)
    assert torch.equal(enc.grad[:, 1:], g[:, :n]) and not enc.grad[:, 0].any()
    assert rel(tok.grad, g[:, n:].sum((0, 1)).view(1, 1, C)) < 1e-5


def test_gather_rows_and_layer_norm_rows():
    torch.manual_seed(1)
    B, G, C, nm = 9, 64, 384, 38
    t = torch.randn(B, G, C, device="cuda")
    order = torch.stack([torch.randperm(G, device="cuda") for _ in range(B)])
    got = ops.gather_rows(t, order, G - nm, nm)
    assert torch.equal(got, torch.gather(t, 1, order[:, G - nm:, None].expand(-1, -1, C)))
    ln = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        ln.weight.normal_(1, 0.1)
        ln.bias.normal_(0, 0.1)
    x = torch.randn(B, G, C, device="cuda", requires_grad=True)
    y = layers.layer_norm_rows(x, ln.weight, ln.bias, ln.eps, G - nm, nm)
    xr = x.detach().clone().requires_grad_(True)
    want = ln(xr[:, -nm:])
    torch.testing.assert_close(y, want, rtol=1e-4, atol=1e-5)
    g = torch.randn_like(want)
    gw, gb = torch.autograd.grad(want, [ln.weight, ln.bias], g, retain_graph=True)
    want.backward(g)
    ln.weight.grad = ln.bias.grad = None
    y.backward(g)
    torch.testing.assert_close(x.grad, xr.grad, rtol=1e-4, atol=1e-5)
    assert not x.grad[:, :G - nm].any()
    torch.testing.assert_close(ln.weight.grad, gw, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ln.bias.grad, gb, rtol=1e-4, atol=1e-4)


def test_drop_path_gates_kernel_distribution():
    keep = torch.tensor([1.0, 1.0, 0.95, 0.95, 0.9, 0.9], device="cuda")
    s1 = torch.tensor([7], dtype=torch.int64, device="cuda")
    s2 = torch.tensor([8], dtype=torch.int64, device="cuda")
    B = 4096
    g1 = ops.drop_path_gates(s1, keep, B, draw_id=1)
    assert torch.equal(g1, ops.drop_path_gates(s1, keep, B, draw_id=1))
    assert not torch.equal(g1, ops.drop_path_gates(s2, keep, B, draw_id=1))
    assert not torch.equal(g1, ops.drop_path_gates(s1, keep, B, draw_id=2))
    assert torch.all(g1[:2] == 1.0)
    for l in range(2, 6):
        kp = keep[l].item()
        vals = g1[l].unique()
        assert all(abs(v) < 1e-6 or abs(v - 1.0 / kp) < 1e-5 for v in vals.tolist())
        assert abs((g1[l] > 0).float().mean().item() - kp) < 0.02        # sigma ~ 0.005
    assert not torch.equal(g1[2], g1[3])                                  # the two branches of a Block draw independently
    # through layers.drop_path_gates with a staged seed (the engine's path) and without (torch.randint seed)
    with layers.drop_path_seed(s1):
        a = layers.drop_path_gates([0.0, 0.1], 64, torch.device("cuda"), True)
    b = layers.drop_path_gates([0.0, 0.1], 64, torch.device("cuda"), True)
    assert a.shape == (4, 64) and b.shape == (4, 64) and torch.all(a[:2] == 1.0)
    assert layers.drop_path_gates([0.0, 0.0], 64, torch.device("cuda"), True) is None


def test_embedding_and_small_helpers():
    table = torch.randn(8192, 384, device="cuda").bfloat16()
    lab = torch.randint(0, 8192, (8192,), device="cuda", dtype=torch.int32)
    assert torch.equal(ops.embedding_bf16(table, lab), table[lab.long()])
    a, b = torch.randn(512, device="cuda"), torch.randn(512, device="cuda")
    want = a + b
    assert torch.equal(ops.accumulate_(a, b), want)
    x = torch.randn(1000, device="cuda")
    want = x * 0.25
    assert torch.equal(ops.scale_by_(x, torch.tensor([0.25], device="cuda")), want)
    z = torch.ones(1 << 20, device="cuda")
    assert not ops.zero_(z).any()
