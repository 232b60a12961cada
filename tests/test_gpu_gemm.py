"""GPU numerics: the tcgen05/TMA GEMM and its fused epilogues vs a plain PyTorch fp32 reference on the SAME
bf16-rounded operands.  Tolerance: fp32 accumulation of exact bf16 products, so the only differences are
summation order (<= 1e-5 relative to the row scale) and, for bf16 outputs, the final rounding (2^-8)."""
import pytest
import torch
import torch.nn.functional as F

from act_b200 import ops

pytestmark = pytest.mark.gpu


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(torch.bfloat16)


def check(got, want, out_bf16):
    want = want.float()
    scale = want.abs().max().item() + 1e-6
    tol = (2 ** -7 if out_bf16 else 2e-5) * scale
    err = (got.float() - want).abs().max().item()
    assert err <= tol, (err, tol)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (3456, 1152, 384), (3456, 384, 1536), (108, 384, 384),
                                   (8192, 1536, 384), (300, 72, 200), (128, 64, 1024), (4096, 512, 512)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_layouts(M, N, K, a_mn, b_mn):
    torch.manual_seed(M + N + K)
    if (a_mn and M % 8) or (b_mn and N % 8) or K % 8:
        pytest.skip("pitch must be a multiple of 8 elements")
    a, b = rnd(M, K), rnd(N, K)
    want = a.float() @ b.float().t()
    A = a.t().contiguous() if a_mn else a
    B = b.t().contiguous() if b_mn else b
    for out_dtype in (torch.float32, torch.bfloat16):
        got = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=out_dtype)
        check(got, want, out_dtype == torch.bfloat16)


@pytest.mark.parametrize("block_n", [64, 128, 192])
def test_gemm_epilogues(block_n):
    torch.manual_seed(0)
    M, N, K = 1000, 384, 256
    a, w = rnd(M, K), rnd(N, K, scale=0.1)
    bias = torch.randn(N, device="cuda")
    lin = F.linear(a.float(), w.float(), bias)
    # bias + GELU with pre-activation side output (fc1)
    pre = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    got = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, preact_out=pre, block_n=block_n)
    check(pre, lin, True)
    check(got, F.gelu(lin), True)
    # bias + ReLU
    check(ops.gemm(a, w, bias=bias, act=ops.ACT_RELU, block_n=block_n), F.relu(lin), True)
    # bias + residual, fp32, in place on the residual stream (proj / fc2)
    x = torch.randn(M, N, device="cuda")
    want = x + lin
    got = ops.gemm(a, w, bias=bias, resid=x, out=x, block_n=block_n)
    assert got.data_ptr() == x.data_ptr()
    check(x, want, False)
    # dgrad through GELU: (dY . W) * gelu'(u), MN-major weight
    dy, u = rnd(M, N), rnd(M, K)
    uu = u.float().requires_grad_(True)
    F.gelu(uu).backward(dy.float() @ w.float())
    got = ops.gemm(dy, w, b_mn=True, mul_in=u, mul_mode=ops.MUL_GELU_GRAD, block_n=block_n)
    check(got, uu.grad, True)
    got = ops.gemm(dy, w, b_mn=True, mul_in=u, mul_mode=ops.MUL_RELU_MASK, block_n=block_n)
    check(got, (dy.float() @ w.float()) * (u.float() > 0), True)
    # alpha
    check(ops.gemm(a, w, alpha=0.125, out_dtype=torch.float32, block_n=block_n), 0.125 * (lin - bias), False)


@pytest.mark.parametrize("splits", [1, 2, 7, 64])
def test_gemm_wgrad_splitk(splits):
    """dW[N,K] = dY^T X with both operands MN-major, split over the token dimension, atomically accumulated."""
    torch.manual_seed(1)
    T, N, K = 3456, 384, 1536
    dy, x = rnd(T, N, scale=0.1), rnd(T, K)
    want = dy.float().t() @ x.float()
    got = ops.gemm(dy, x, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=splits)
    check(got, want, False)
    acc = torch.ones(N, K, device="cuda")
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=acc, splits=max(splits, 2))
    check(acc, want + 1, False)


def test_gemm_strided_views():
    """Operands / outputs that are column slices of wider buffers (pitch != width), as the qkv split uses."""
    torch.manual_seed(2)
    big = rnd(512, 1152)
    w = rnd(384, 384, scale=0.1)
    a = big[:, 384:768]
    out_big = torch.zeros(512, 1024, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, w, out=out_big[:, 128:512])
    check(out_big[:, 128:512], a.float() @ w.float().t(), True)
    assert (out_big[:, :128] == 0).all() and (out_big[:, 512:] == 0).all()


@pytest.mark.parametrize("M,N,K,bn", [(262144, 256, 128, 0), (65536, 512, 256, 256), (65536, 384, 512, 128),
                                      (40000, 512, 512, 256), (128 * 700, 128, 64, 128)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_gemm_persistent(M, N, K, bn, a_mn, b_mn):
    """Persistent kernel (one CTA per SM, double-buffered TMEM accumulator) vs the PyTorch reference and vs the
    one-tile-per-CTA kernel."""
    torch.manual_seed(N + K)
    a, b = rnd(M, K), rnd(N, K, scale=0.2)
    bias = torch.randn(N, device="cuda")
    A = a.t().contiguous() if a_mn else a
    B = b.t().contiguous() if b_mn else b
    got = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, bias=bias, act=ops.ACT_RELU, block_n=bn, persistent=1)
    ref = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, bias=bias, act=ops.ACT_RELU, persistent=0)
    assert torch.equal(got, ref)
    sl = slice(0, M, 37)
    check(got[sl], torch.relu(a[sl].float() @ b.float().t() + bias), True)


def test_gemm_persistent_splitk_wgrad():
    torch.manual_seed(5)
    T, N, K = 262144, 384, 512
    dy, x = rnd(T, N, scale=0.05), rnd(T, K)
    want = (dy.double().t() @ x.double()).float()
    scale = want.abs().max().item()
    for persistent in (0, 1):
        got = ops.gemm(dy, x, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=ops.wgrad_splits(N, K, T),
                       persistent=persistent)
        # fp32 accumulation of 262144 / splits exact products per accumulator: 2^-24 * sqrt(terms) ~ 2e-5 relative to the
        # row scale, against an fp64 reference (one CTA per SM since round 2: fewer, longer chains than the 2e-5 of check())
        err = (got - want).abs().max().item()
        assert err <= 1e-4 * scale, (err, scale)


@pytest.mark.parametrize("persistent", [0, 1, -1])            # -1: automatic choice = the CTA-pair 256 x 384 tiles here
def test_gemm_fused_group_max(persistent):
    """Fused max over each 32 consecutive rows on the fp32 accumulators (+ arg-max), with and without `out`."""
    torch.manual_seed(6)
    G, N, K = 4100, 384, 512
    M = G * 32
    a, w = rnd(M, K), rnd(N, K, scale=0.1)
    bias = torch.randn(N, device="cuda")
    full = a.float() @ w.float().t() + bias
    v, i = full.view(G, 32, N).max(1)
    gf = torch.empty(G, N, device="cuda")
    gb = torch.empty(G, N, dtype=torch.bfloat16, device="cuda")
    ga = torch.empty(G, N, dtype=torch.uint8, device="cuda")
    out = ops.gemm(a, w, bias=bias, gmax_f32=gf, gmax_bf16=gb, garg=ga, persistent=persistent)
    check(out, full, True)
    torch.testing.assert_close(gf, v, rtol=1e-4, atol=1e-4)
    assert torch.equal(gb, gf.bfloat16())
    picked = torch.gather(full.view(G, 32, N), 1, ga.long()[:, None]).squeeze(1)
    torch.testing.assert_close(picked, v, rtol=1e-4, atol=1e-4)       # arg points at a (near-)maximal row
    assert (ga.long() == i).float().mean().item() > 0.999
    gf2 = torch.zeros(G, N, device="cuda")
    assert ops.gemm(a, w, bias=bias, gmax_f32=gf2, no_out=True, persistent=persistent) is None
    assert torch.equal(gf2, gf)
    # no output tile, every fused-max output at once: same bits, same arg-max
    gf3, gb3 = torch.zeros(G, N, device="cuda"), torch.zeros(G, N, dtype=torch.bfloat16, device="cuda")
    ga3 = torch.full((G, N), 255, dtype=torch.uint8, device="cuda")
    ops.gemm(a, w, bias=bias, gmax_f32=gf3, gmax_bf16=gb3, garg=ga3, no_out=True, persistent=persistent)
    assert torch.equal(gf3, gf) and torch.equal(gb3, gb) and torch.equal(ga3, ga)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_gemm_bn192(a_mn, b_mn):
    """128 x 192 tiles (N = 1536 / 1152 in one wave of the one-tile-per-CTA kernel)."""
    torch.manual_seed(9)
    M, N, K = 3456, 1536, 384
    a, b = rnd(M, K), rnd(N, K, scale=0.1)
    A = a.t().contiguous() if a_mn else a
    B = b.t().contiguous() if b_mn else b
    got = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, block_n=192, out_dtype=torch.float32)
    check(got, a.float() @ b.float().t(), False)
    auto = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    assert torch.equal(got, auto)          # the heuristic picks the 192-wide tile here


@pytest.mark.parametrize("M,K", [(8192, 768), (8192, 3072), (8100, 768), (6200, 1024)])
def test_gemm_wide384_tiles(M, K):
    """Narrow output, long K, ~one round of 128 x 384 tiles (two 192-wide MMAs sharing the A tile): the teacher ViT's
    proj / fc2 on the token rows.  Plain and bias + residual (fp32, in place) epilogues, bf16 output, ragged M."""
    torch.manual_seed(M + K)
    N = 768
    a, w = rnd(M, K), rnd(N, K, scale=0.05)
    bias = torch.randn(N, device="cuda")
    lin = a.float() @ w.float().t()
    check(ops.gemm(a, w, out_dtype=torch.float32), lin, False)
    check(ops.gemm(a, w, bias=bias), lin + bias, True)
    x = torch.randn(M, N, device="cuda")
    want = x + lin + bias
    got = ops.gemm(a, w, bias=bias, resid=x, out_dtype=torch.float32)
    check(got, want, False)
    ops.gemm(a, w, bias=bias, resid=x, out=x)          # in place on the residual stream
    check(x, want, False)
    # an epilogue the wide configuration does not instantiate falls back to the regular tiles
    check(ops.gemm(a, w, bias=bias, act=ops.ACT_GELU), F.gelu(lin + bias), True)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (8192, 768, 768), (8100, 2304, 768), (300, 264, 200), (4096, 512, 1024)])
def test_gemm_cta_pair(M, N, K):
    """The cta_group::2 kernel (two CTAs of a cluster on one 256 x 256 tile), forced with persistent=2: plain, GELU and
    bias + residual epilogues, ragged M / N, against fp32 torch on the same bf16 operands."""
    torch.manual_seed(M + N + K)
    a, w = rnd(M, K), rnd(N, K, scale=0.05)
    bias = torch.randn(N, device="cuda")
    lin = a.float() @ w.float().t()
    check(ops.gemm(a, w, out_dtype=torch.float32, persistent=2), lin, False)
    check(ops.gemm(a, w, bias=bias, persistent=2), lin + bias, True)
    # bf16 output through TMA (plain epilogue): a column slice of a wider buffer -- the tensor map must clip at N, not at
    # the 64-column box, and leave the neighbouring columns alone
    wide = torch.zeros(M, N + 192, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, w, bias=bias, out=wide[:, 64:64 + N], persistent=2)
    check(wide[:, 64:64 + N], lin + bias, True)
    assert (wide[:, :64] == 0).all() and (wide[:, 64 + N:] == 0).all()
    check(ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, persistent=2), F.gelu(lin + bias), True)
    x = torch.randn(M, N, device="cuda")
    want = x + lin + bias
    ops.gemm(a, w, bias=bias, resid=x, out=x, persistent=2)          # in place on the residual stream
    check(x, want, False)
    res = torch.randn((M + 31) // 32, N, device="cuda")              # per-group broadcast term (mini-PointNet conv3)
    if M % 32 == 0:
        got = ops.gemm(a, w, resid=res, resid_row_div=32, persistent=2)
        check(got, lin + res.repeat_interleave(32, 0), True)


@pytest.mark.parametrize("M,persist", [(32768, -1), (4096, -1), (262144, -1), (2048, 1)])
def test_gemm_fused_column_statistics(M, persist):
    """colstats: per-column sum / sum of squares of the stored values (accumulator + bias + per-group term), accumulated per
    CTA in shared memory and flushed once -- the BatchNorm statistics of the mini-PointNet's conv3 (models/dvae.py:196-197)
    without a second pass over the [M,512] output.  Pair kernel (large M) and persistent kernel (small M)."""
    import torch
    from act_b200 import ops
    torch.manual_seed(M)
    N, K, k = 512, 256, 32
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    gpart = torch.randn(M // k, N, device="cuda")
    stats = torch.empty(2, N, device="cuda")
    out = ops.gemm(a, w, resid=gpart, resid_row_div=k, colstats=stats, persistent=persist)
    ref = a.float() @ w.float().t() + gpart.repeat_interleave(k, 0)
    assert ((out.float() - ref).norm() / ref.norm()).item() < 5e-3
    torch.testing.assert_close(stats[0], ref.sum(0), rtol=2e-3, atol=2e-2 * M ** 0.5)
    torch.testing.assert_close(stats[1], (ref * ref).sum(0), rtol=2e-3, atol=1e-3)
    mean, var = stats[0] / M, stats[1] / M - (stats[0] / M) ** 2
    torch.testing.assert_close(mean, ref.mean(0), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(var, ref.var(0, unbiased=False), rtol=2e-3, atol=1e-5)
