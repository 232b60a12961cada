"""GPU numerics of the non-GEMM kernels and of the fused autograd regions vs plain PyTorch fp32 references.
Tolerances are stated per test: fp32 kernels (LayerNorm, loss, AdamW) agree to ~1e-5; anything that passes
through a bf16 GEMM operand carries bf16 rounding (2^-8 relative per element), so fused regions are compared
with a relative Frobenius-norm error bound."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from act_b200 import layers, modules, ops
from oracle import ref_model

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.mark.parametrize("M,C", [(108, 384), (3456, 384), (1000, 768), (64, 128)])
def test_layernorm_fwd_bwd(M, C):
    torch.manual_seed(0)
    x = torch.randn(M, C, device="cuda") * 2 + 0.5
    pos = torch.randn(M, C, device="cuda")
    w = torch.randn(C, device="cuda") * 0.2 + 1
    b = torch.randn(C, device="cuda") * 0.1
    xs_ref = (x + pos).requires_grad_(True)
    y_ref = F.layer_norm(xs_ref, (C,), w.clone().requires_grad_(True), b, 1e-5)
    y, xs, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5, pos=pos, out_dtype=torch.float32)
    torch.testing.assert_close(xs, x + pos)
    torch.testing.assert_close(y, y_ref, rtol=1e-5, atol=1e-5)
    yb, _, _, _ = ops.layernorm_fwd(x, w, b, 1e-5, pos=pos)
    assert yb.dtype == torch.bfloat16 and rel(yb, y_ref) < 4e-3
    # backward (fp32 dy), with a residual-branch gradient, a pos accumulator, gated bf16 copy and bias sums
    dy = torch.randn(M, C, device="cuda")
    dres = torch.randn(M, C, device="cuda")
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    xr = (x + pos).clone().requires_grad_(True)
    F.layer_norm(xr, (C,), wr, br, 1e-5).backward(dy)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dacc = torch.ones(M, C, device="cuda")
    T = 4 if M % 4 == 0 else 1
    gate = torch.rand(M // T, device="cuda")
    dbias = torch.zeros(C, device="cuda")
    dx, g = ops.layernorm_bwd(dy, xs, mean, rstd, w, dg, db, dres=dres, dacc=dacc, want_bf16=True, row_scale=gate,
                              rows_per_scale=T, dbias=dbias)
    want = xr.grad + dres
    torch.testing.assert_close(dx, want, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dacc, want + 1, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dg, wr.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(db, br.grad, rtol=1e-3, atol=1e-3)
    gated = want * gate.repeat_interleave(T)[:, None]
    assert rel(g, gated) < 4e-3
    torch.testing.assert_close(dbias, gated.sum(0), rtol=1e-3, atol=1e-3)
    # bf16 dy path
    dx2, _ = ops.layernorm_bwd(dy.bfloat16(), xs, mean, rstd, w, torch.zeros_like(dg), torch.zeros_like(db))
    assert rel(dx2, xr.grad) < 5e-3
    g2 = ops.cast_rows(want, gate, T, dbias=(db2 := torch.zeros(C, device="cuda")))
    assert rel(g2, gated) < 4e-3
    torch.testing.assert_close(db2, gated.sum(0), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("B,T,H", [(3, 27, 6), (2, 64, 6), (2, 14, 6), (1, 65, 6), (2, 206, 6), (1, 130, 12)])
def test_attention_fwd_bwd(B, T, H):
    torch.manual_seed(T)
    C = H * 64
    qkv = (torch.randn(B * T, 3 * C, device="cuda") * 0.7).bfloat16()
    scale = 0.125
    r = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4).requires_grad_(True)
    q, k, v = r[0], r[1], r[2]
    attn = ((q @ k.transpose(-2, -1)) * scale).softmax(-1)
    o_ref = (attn @ v).transpose(1, 2).reshape(B * T, C)
    o, lse = ops.attention_fwd(qkv, B, T, H, scale)
    assert rel(o, o_ref) < 4e-3
    torch.testing.assert_close(lse, torch.logsumexp((q @ k.transpose(-2, -1)) * scale, -1), rtol=1e-4, atol=1e-4)
    do = (torch.randn(B * T, C, device="cuda") * 0.5).bfloat16()
    o_ref.backward(do.float())
    dqkv_ref = r.grad.permute(1, 3, 0, 2, 4).reshape(B * T, 3 * C)
    dqkv = ops.attention_bwd(qkv, o, do, lse, B, T, H, scale)
    assert rel(dqkv, dqkv_ref) < 1e-2


@pytest.mark.parametrize("M,C,dt", [(3456, 1536, torch.bfloat16), (3456, 384, torch.bfloat16), (8192, 1152, torch.bfloat16),
                                    (1003, 2048, torch.bfloat16), (37, 8, torch.bfloat16), (4864, 384, torch.float32),
                                    (70000, 512, torch.bfloat16)])
def test_colsum_token_sized_and_large(M, C, dt):
    """Bias-gradient column sums: the (column block x row split) kernel for token-sized matrices and the chan_reduce
    kernel for the [B*G*k, C] activations (M >= 65536) accumulate into `out` and agree with an fp64 sum."""
    torch.manual_seed(M + C)
    x = torch.randn(M, C, device="cuda").to(dt)
    out = torch.full((C,), 0.5, device="cuda")
    ops.colsum(x, out)
    want = x.double().sum(0) + 0.5
    torch.testing.assert_close(out.double(), want, rtol=1e-4, atol=2e-3 * M ** 0.5)


def test_colsum_loss_adamw():
    torch.manual_seed(1)
    x = torch.randn(1000, 1536, device="cuda")
    out = torch.ones(1536, device="cuda")
    ops.colsum(x.bfloat16(), out)
    torch.testing.assert_close(out, x.bfloat16().float().sum(0) + 1, rtol=1e-3, atol=1e-2)
    out = torch.zeros(200, device="cuda")
    ops.colsum(x[:, 100:300], out)
    torch.testing.assert_close(out, x[:, 100:300].sum(0), rtol=1e-4, atol=1e-3)
    # cosine loss == act.py:1243-1254 loop
    B, nm, C = 4, 38, 384
    s = torch.randn(B, nm, C, device="cuda", requires_grad=True)
    t = torch.randn(B, nm, C, device="cuda")
    ref = sum(1 - F.cosine_similarity(s[b], t[b], 1, 1e-8).mean() for b in range(B)) / B
    ref.backward()
    s2 = s.detach().clone().requires_grad_(True)
    loss = layers.cosine_loss(s2, t)
    (loss * 2.0).backward()
    torch.testing.assert_close(loss, ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(s2.grad, 2 * s.grad, rtol=1e-4, atol=1e-7)
    # AdamW == torch.optim.AdamW for 3 steps, decay + no-decay groups
    lin = torch.nn.Linear(40, 24).cuda()
    ref_lin = torch.nn.Linear(40, 24).cuda()
    ref_lin.load_state_dict(lin.state_dict())
    opt = torch.optim.AdamW([{"params": [ref_lin.weight], "weight_decay": 0.05},
                             {"params": [ref_lin.bias], "weight_decay": 0.0}], lr=1e-2)
    fp = layers.FlatParams(lin, lr=1e-2, weight_decay=0.05)
    for step in range(3):
        g = torch.Generator(device="cuda").manual_seed(step)
        gw, gb = torch.randn(24, 40, device="cuda", generator=g), torch.randn(24, device="cuda", generator=g)
        ref_lin.weight.grad, ref_lin.bias.grad = gw.clone(), gb.clone()
        opt.step()
        lin.weight.grad.copy_(gw)
        lin.bias.grad.copy_(gb)
        fp.set_hyper()
        fp.step()
    torch.testing.assert_close(lin.weight, ref_lin.weight, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(lin.bias, ref_lin.bias, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(lin.weight._act_shadow.float(), ref_lin.weight, rtol=1e-2, atol=1e-3)


def test_pointnet_pieces():
    torch.manual_seed(2)
    G, k, C = 50, 32, 256
    x = torch.randn(G * k, C, device="cuda").bfloat16()
    ob, of, arg = ops.group_max(x, k, want_f32=True)
    v, i = x.float().view(G, k, C).max(1)
    assert torch.equal(of, v) and torch.equal(ob.float(), v) and torch.equal(arg.long(), i)
    d = torch.randn(G, C, device="cuda")
    dF = ops.group_max_bwd(d, arg, k)
    want = torch.zeros(G, k, C, device="cuda").scatter_(1, i[:, None], d.bfloat16().float()[:, None])
    assert torch.equal(dF.float().view(G, k, C), want)
    dF2 = ops.group_max_bwd(d, arg, k, out=dF.clone())
    assert rel(dF2.float().view(G, k, C), 2 * want) < 4e-3
    sb, sf = ops.group_sum(x, k, want_f32=True)
    torch.testing.assert_close(sf, x.float().view(G, k, C).sum(1), rtol=1e-5, atol=1e-4)
    # BN stats / apply / backward vs torch batch_norm on the same bf16-rounded input
    M, C = 4096, 512
    h = (torch.randn(M, C, device="cuda") * 1.5 + 0.3).bfloat16()
    gam, bet = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
    s1, s2 = ops.bn_stats(h)
    mean = s1 / M
    var = s2 / M - mean * mean
    hf = h.float().requires_grad_(True)
    gr = gam.clone().requires_grad_(True)
    br = bet.clone().requires_grad_(True)
    y_ref = F.relu(F.batch_norm(hf, None, None, gr, br, True, 0.1, 1e-5))
    torch.testing.assert_close(mean, hf.detach().mean(0), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(var, hf.detach().var(0, unbiased=False), rtol=1e-3, atol=1e-4)
    rstd = torch.rsqrt(var + 1e-5)
    y = ops.bn_apply(h, gam * rstd, bet - mean * gam * rstd, relu=True)
    assert rel(y, y_ref) < 4e-3
    dy = torch.randn(M, C, device="cuda")
    y_ref.backward(dy)
    dz = (dy * (y_ref > 0)).bfloat16()
    dh, dbeta, dgamma = ops.bn_bwd(dz, h, mean, rstd, gam)
    assert rel(dh, hf.grad) < 1e-2
    assert rel(dbeta, br.grad) < 5e-3 and rel(dgamma, gr.grad) < 5e-3
    # conv1 + analytic BN1 statistics
    p = torch.randn(M, 3, device="cuda") * 0.1
    W, b = torch.randn(128, 3, device="cuda"), torch.randn(128, device="cuda") * 0.1
    mom = ops.pn_moments(p) / M
    torch.testing.assert_close(mom[:3].float(), p.mean(0), rtol=1e-5, atol=1e-7)
    a1 = ops.pn_conv1(p, W, b, relu=True)
    assert rel(a1, F.relu(p @ W.t() + b)) < 4e-3


def test_block_stack_vs_oracle_cfg1(golden):
    """BASELINE config 1 (Block x12, d=384, 64 tokens, batch 2, eval) against the golden output of the
    reference's utils/transformer_layers.Block.  bf16 tensor-core operands: relative Frobenius error < 1e-2."""
    g = golden("block12_cfg1.npz")
    blocks = torch.nn.ModuleList([modules.Block(384, 6) for _ in range(12)])
    ref_model.fill_params(blocks, seed=1)
    blocks = blocks.cuda().eval()
    x = torch.from_numpy(g["x"]).cuda()
    with torch.no_grad():
        y = modules.run_blocks(list(blocks), x, None, False)
        y1 = x
        for b in blocks:
            y1 = b(y1)
    want = torch.from_numpy(g["y"]).cuda()
    assert rel(y, want) < 1e-2, rel(y, want)
    assert rel(y1, want) < 1e-2


def test_block_stack_backward_vs_torch():
    """2 Blocks with pos and DropPath gates: output, input/pos gradients and every parameter gradient vs the
    oracle's plain-PyTorch Block (fp32) driven with the same gates."""
    torch.manual_seed(3)
    B, T, C = 8, 27, 384
    mine = torch.nn.ModuleList([modules.Block(C, 6) for _ in range(2)])
    ref_model.fill_params(mine, seed=7)
    ref = torch.nn.ModuleList([ref_model.Block(C, 6) for _ in range(2)])
    ref.load_state_dict(mine.state_dict())
    mine, ref = mine.cuda().train(), ref.cuda().train()
    x = torch.randn(B, T, C, device="cuda", requires_grad=True)
    pos = (torch.randn(B, T, C, device="cuda") * 0.3).requires_grad_(True)
    gates = (torch.rand(4, B, device="cuda") > 0.3).float() / 0.7
    y = layers.transformer_stack(x, pos, list(mine), 6, 1e-5, gates)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr, pr = x.detach().clone().requires_grad_(True), pos.detach().clone().requires_grad_(True)
    cur = xr
    for l, blk in enumerate(ref):
        cur = cur + pr
        cur = cur + gates[2 * l].view(B, 1, 1) * blk.attn(blk.norm1(cur))
        cur = cur + gates[2 * l + 1].view(B, 1, 1) * blk.mlp(blk.norm2(cur))
    cur.backward(dy)
    assert rel(y, cur) < 5e-3
    assert rel(x.grad, xr.grad) < 2e-2 and rel(pos.grad, pr.grad) < 2e-2
    for (n, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        assert rel(p.grad, q.grad) < 3e-2, (n, rel(p.grad, q.grad))


def test_linear_fn():
    torch.manual_seed(4)
    lin = torch.nn.Linear(128, 384).cuda()
    x = torch.randn(5, 26, 128, device="cuda", requires_grad=True)
    for gelu in (False, True):
        lin.zero_grad()
        x.grad = None
        y = layers.linear(x, lin.weight, lin.bias, gelu=gelu)
        dy = torch.randn_like(y)
        y.backward(dy)
        got = (x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
        lin.zero_grad()
        x.grad = None
        yr = F.linear(x, lin.weight, lin.bias)
        yr = F.gelu(yr) if gelu else yr
        yr.backward(dy)
        assert rel(y, yr) < 5e-3
        for a, b in zip(got, (x.grad, lin.weight.grad, lin.bias.grad)):
            assert rel(a, b) < 1e-2
