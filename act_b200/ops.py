"""Tensor-level wrappers over the C ABI (include/act_b200.h).  Every function here launches hand-written
sm_100a kernels from libact_b200.so on the current CUDA stream; none has a PyTorch/CPU fallback."""
import torch

from . import _lib

LAUNCHES = 0  # number of libact_b200 kernel launches issued through this module (bench.py reads it)
import os as _os

# Precision of the dense path.  "bf16" (default, the speed mode): GEMM operands and stored activations are bf16, fp32
# accumulation.  "fp32x3" (the PARITY mode, north_star's 1e-3 bar): activations are stored in f32 and every GEMM operand
# is split into bf16 (hi, mid, lo) pieces so that the same tcgen05 kernel computes hi hi + hi mid + mid hi (+ mid mid +
# hi lo + lo hi, PARITY_TERMS = 6, the default) over a 3- / 6-fold K (csrc/parity.cu) -- ~16 / ~24 mantissa bits per
# product, fp32 accumulation; attention runs in f32 on the FMA pipes.
_ACT_DTYPE = torch.float32 if _os.environ.get("ACT_B200_PRECISION", "bf16") == "fp32x3" else torch.bfloat16


def act_dtype():
    """Storage dtype of the activations that feed GEMMs: bf16 (speed mode), f32 (parity mode)."""
    return _ACT_DTYPE


def set_precision(mode):
    global _ACT_DTYPE
    if mode not in ("bf16", "fp32x3"):
        raise ValueError("precision: 'bf16' or 'fp32x3'")
    _ACT_DTYPE = torch.float32 if mode == "fp32x3" else torch.bfloat16


def get_precision():
    return "fp32x3" if _ACT_DTYPE == torch.float32 else "bf16"


class precision:
    """with ops.precision("fp32x3"): ...   (modules built / called inside run the parity mode)"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = get_precision()
        set_precision(self.mode)

    def __exit__(self, *a):
        set_precision(self.prev)


class gemm_sm_cap:
    """with ops.gemm_sm_cap(n): many-tile GEMMs ENQUEUED inside occupy at most n SMs (0 = all).  Launch-time host state."""

    def __init__(self, n):
        self.n = int(n or 0)

    def __enter__(self):
        if self.n:
            _lib.check(_lib.lib().act_set_option(3, self.n), "act_set_option")

    def __exit__(self, *a):
        if self.n:
            _lib.check(_lib.lib().act_set_option(3, 0), "act_set_option")


def _io32(t):
    return int(t.dtype == torch.float32)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------ Group tokenizer
def furthest_point_sample(xyz, npoint, return_center=False):
    """pointnet2_utils.furthest_point_sample (utils/misc.py:44): xyz [B,N,3] f32 -> idx [B,npoint] i32."""
    xyz = _f32c(xyz)
    B, N, _ = xyz.shape
    idx = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    center = torch.empty(B, npoint, 3, dtype=torch.float32, device=xyz.device) if return_center else None
    _lib.call("act_fps", xyz, B, N, npoint, idx, center)
    _count()
    return (idx, center) if return_center else idx


class _GatherOperation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        features = _f32c(features)
        idx = idx.to(torch.int32).contiguous()
        B, C, N = features.shape
        M = idx.shape[1]
        out = torch.empty(B, C, M, dtype=torch.float32, device=features.device)
        _lib.call("act_gather_points", features, idx, B, C, N, M, out)
        _count()
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    def backward(ctx, gout):
        (idx,) = ctx.saved_tensors
        gout = _f32c(gout)
        B, C, M = gout.shape
        gfeat = torch.empty(B, C, ctx.N, dtype=torch.float32, device=gout.device)
        _lib.call("act_gather_points_grad", gout, idx, B, C, ctx.N, M, gfeat)
        _count(2)
        return gfeat, None


gather_operation = _GatherOperation.apply


@torch.no_grad()
def knn(ref, query, k, want_dist=True, want_neighborhood=False):
    """knn_cuda.KNN(k, transpose_mode=True): ref [B,N,3], query [B,Q,3] -> (dist [B,Q,k] f32, idx [B,Q,k] i64
    [, neighborhood [B,Q,k,3] = ref[idx] - query])."""
    ref, query = _f32c(ref), _f32c(query)
    B, N, _ = ref.shape
    Q = query.shape[1]
    idx = torch.empty(B, Q, k, dtype=torch.int64, device=ref.device)
    dist = torch.empty(B, Q, k, dtype=torch.float32, device=ref.device) if want_dist else None
    nb = torch.empty(B, Q, k, 3, dtype=torch.float32, device=ref.device) if want_neighborhood else None
    _lib.call("act_knn", ref, query, B, N, Q, k, dist, idx, nb)
    _count()
    return dist, idx, nb


@torch.no_grad()
def group(xyz, num_group, group_size):
    """Group.forward (models/dvae.py:161-183): xyz [B,N,3] -> (neighborhood [B,G,k,3], center [B,G,3],
    idx [B,G,k] i64, fps_idx [B,G] i32) in two launches."""
    xyz = _f32c(xyz)
    B, N, _ = xyz.shape
    dev = xyz.device
    fps_idx = torch.empty(B, num_group, dtype=torch.int32, device=dev)
    center = torch.empty(B, num_group, 3, dtype=torch.float32, device=dev)
    idx = torch.empty(B, num_group, group_size, dtype=torch.int64, device=dev)
    nb = torch.empty(B, num_group, group_size, 3, dtype=torch.float32, device=dev)
    _lib.call("act_group", xyz, B, N, num_group, group_size, fps_idx, center, idx, nb)
    _count(2)
    return nb, center, idx, fps_idx


# ------------------------------------------------------------------------------------------- Chamfer
def chamfer_forward(xyz1, xyz2):
    """chamfer.forward (extensions/chamfer_dist/chamfer_cuda.cpp:22-25)."""
    xyz1, xyz2 = _f32c(xyz1), _f32c(xyz2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    d1 = torch.empty(B, n, dtype=torch.float32, device=dev)
    d2 = torch.empty(B, m, dtype=torch.float32, device=dev)
    i1 = torch.empty(B, n, dtype=torch.int32, device=dev)
    i2 = torch.empty(B, m, dtype=torch.int32, device=dev)
    _lib.call("act_chamfer_forward", xyz1, xyz2, B, n, m, d1, d2, i1, i2)
    _count()
    return [d1, d2, i1, i2]


def chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    """chamfer.backward (chamfer_cuda.cpp:27-34)."""
    xyz1, xyz2 = _f32c(xyz1), _f32c(xyz2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = torch.empty_like(xyz1)
    gx2 = torch.empty_like(xyz2)
    _lib.call("act_chamfer_backward", xyz1, xyz2, idx1.contiguous(), idx2.contiguous(), _f32c(grad_dist1),
              _f32c(grad_dist2), B, n, m, gx1, gx2)
    _count(3)
    return [gx1, gx2]


class ChamferFunction(torch.autograd.Function):
    """extensions/chamfer_dist/__init__.py:13-25."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        d1, d2, i1, i2 = chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, i1, i2)
        ctx.mark_non_differentiable(i1, i2)
        return d1, d2

    @staticmethod
    def backward(ctx, g1, g2):
        xyz1, xyz2, i1, i2 = ctx.saved_tensors
        gx1, gx2 = chamfer_backward(xyz1, xyz2, i1, i2, g1, g2)
        return gx1, gx2


# ------------------------------------------------------------------------------ tcgen05 GEMM + epilogues
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
MUL_NONE, MUL_GELU_GRAD, MUL_RELU_MASK = 0, 1, 2
_vp = _lib.ctypes.c_void_p


def _p(t):
    return _vp(t.data_ptr()) if t is not None else None


# product terms of the parity-mode GEMM: 6 (hh + hm + mh + mm + hl + lh: ~24 mantissa bits, fp32 grade; default) or 3
# (the first three: ~16 bits, half the work)
PARITY_TERMS = int(_os.environ.get("ACT_B200_PARITY_TERMS", "6"))


def split3(x, mn_major, role_b, pieces=None):
    """f32 2-D operand (last dim contiguous) -> its multi-piece bf16 form for the parity-mode GEMM (csrc/parity.cu)."""
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    P = pieces or PARITY_TERMS
    R, Cc = x.shape
    out = torch.empty((P * R, Cc) if mn_major else (R, P * Cc), dtype=torch.bfloat16, device=x.device)
    _lib.call("act_split_bf16", _p(x), _lib.ctypes.c_int64(R), Cc, _lib.ctypes.c_int64(x.stride(0)), int(mn_major),
              int(role_b), P, out)
    _count()
    return out


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=None, bias=None, act=ACT_NONE,
         preact_out=None, mul_in=None, mul_mode=MUL_NONE, resid=None, resid_row_div=1, row_scale=None,
         rows_per_scale=1, gmax_f32=None, gmax_bf16=None, garg=None, no_out=False, alpha=1.0, splits=1, block_n=0,
         persistent=-1, colstats=None):
    """out[M,N] = epilogue(alpha * A . B^T) on the tcgen05 GEMM (include/act_b200.h: act_gemm_bf16).
    colstats: f32 [2, N] receiving the per-column (sum, sum of squares) of the stored values (BatchNorm statistics).
    a: [M,K] (or [K,M] if a_mn);  b: [N,K] (or [K,N] if b_mn).  2-D, last-dim contiguous; bf16, or f32 in the parity mode
    (both operands are then split into bf16 pieces and the product runs over 3K, see split3).
    out_dtype None = the activation dtype of the current precision mode.
    gmax_f32 / gmax_bf16 / garg: [M/32, N] outputs of the fused per-32-row max; no_out=True skips `out`."""
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    K, M = (a.shape if a_mn else a.shape[::-1])
    Kb, N = (b.shape if b_mn else b.shape[::-1])
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    if a.dtype == torch.float32 or b.dtype == torch.float32:
        a = split3(a.float() if a.dtype != torch.float32 else a, a_mn, 0)
        b = split3(b.float() if b.dtype != torch.float32 else b, b_mn, 1)
        K = PARITY_TERMS * K
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    if out_dtype is None:
        out_dtype = act_dtype()
    if no_out:
        out = None
    elif out is None:
        out = (torch.zeros if splits > 1 else torch.empty)(M, N, dtype=out_dtype, device=a.device)
    if out is not None:
        assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    aux = [t for t in (preact_out, mul_in) if t is not None]
    aux32 = bool(aux) and aux[0].dtype == torch.float32
    assert all((t.dtype == torch.float32) == aux32 for t in aux)
    if preact_out is not None:
        assert preact_out.dtype in (torch.bfloat16, torch.float32) and preact_out.stride(0) == out.stride(0)
    gm = [t for t in (gmax_f32, gmax_bf16, garg) if t is not None]
    ldg = gm[0].stride(0) if gm else 0
    assert all(t.stride(0) == ldg and t.shape == (M // 32, N) for t in gm)
    _lib.call("act_gemm_bf16", _p(a), _p(b), M, N, K, int(a_mn), int(b_mn), a.stride(0), b.stride(0), _p(out),
              out.stride(0) if out is not None else 0, int(out is not None and out.dtype == torch.float32), bias,
              int(act), _p(preact_out), _p(mul_in), mul_in.stride(0) if mul_in is not None else 0, int(mul_mode),
              _p(resid), resid.stride(0) if resid is not None else 0, int(resid_row_div), row_scale,
              int(rows_per_scale), gmax_f32, gmax_bf16, garg, int(ldg), float(alpha), int(splits), int(block_n),
              int(persistent), int(aux32), colstats[0] if colstats is not None else None,
              colstats[1] if colstats is not None else None)
    _count()
    return out


_WGRAD_CTAS = int(_os.environ.get("ACT_B200_WGRAD_CTAS", "148"))   # measured: 148 beats 296 / 222 / 111 / 74 (fewer split-K atomics)


def wgrad_splits(n_out, k_out, tokens):
    """Split-K factor for dW[n_out,k_out] = dY^T X over `tokens`: aim at one CTA per SM -- every split adds a full tile of
    fp32 atomics, and the step is bound by total SM-time (sweep: 296 CTAs 7.43 ms, 222 7.37, 148 7.30-7.34, 111 7.35, 74 7.39)."""
    tiles = ((n_out + 127) // 128) * ((k_out + 127) // 128)
    kb = (tokens + 63) // 64
    return max(1, min(kb, (_WGRAD_CTAS + tiles - 1) // tiles))


def wgrad(dy, x, grad_out):
    """grad_out[N,K] (f32, accumulated in place) += dy[T,N]^T . x[T,K]   (both operands read MN-major)."""
    T, N = dy.shape
    K = x.shape[1]
    sp = wgrad_splits(N, K, T)
    if sp == 1:
        gemm(dy, x, a_mn=True, b_mn=True, out=grad_out, resid=grad_out)
    else:
        gemm(dy, x, a_mn=True, b_mn=True, out=grad_out, splits=sp)
    return grad_out


# ------------------------------------------------------------------------------ Block pieces
def layernorm_fwd(x, gamma, beta, eps=1e-5, pos=None, out_dtype=None, want_sum=False, save_stats=True):
    """x f32 [M,C] (+ pos) -> (y, xsum or None, mean, rstd)."""
    M, C = x.shape
    if out_dtype is None:
        out_dtype = act_dtype()
    y = torch.empty(M, C, dtype=out_dtype, device=x.device)
    xs = torch.empty_like(x) if (want_sum or pos is not None) else None
    mean = torch.empty(M, dtype=torch.float32, device=x.device) if save_stats else None
    rstd = torch.empty(M, dtype=torch.float32, device=x.device) if save_stats else None
    _lib.call("act_layernorm_fwd", x, pos, gamma, beta, float(eps), M, C, xs, _p(y), int(out_dtype == torch.float32),
              mean, rstd)
    _count()
    return y, xs, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, dres=None, dacc=None, want_bf16=False, row_scale=None,
                  rows_per_scale=1, dbias=None):
    """-> (dx f32 [M,C], g (activation dtype) or None); dgamma/dbeta/dbias/dacc accumulated in place."""
    M, C = x.shape
    dx = torch.empty(M, C, dtype=torch.float32, device=x.device)
    g = torch.empty(M, C, dtype=act_dtype(), device=x.device) if want_bf16 else None
    _lib.call("act_layernorm_bwd", _p(dy), int(dy.dtype == torch.float32), x, mean, rstd, gamma, dres, M, C, dx,
              dgamma, dbeta, dacc, _p(g), int(g is not None and g.dtype == torch.float32), row_scale,
              int(rows_per_scale), dbias)
    _count()
    return dx, g


def cast_rows(x, row_scale=None, rows_per_scale=1, dbias=None):
    M, C = x.shape
    g = torch.empty(M, C, dtype=act_dtype(), device=x.device)
    _lib.call("act_cast_rows", x, M, C, row_scale, int(rows_per_scale), _p(g), _io32(g), dbias)
    _count()
    return g


def to_act(x):
    """x [M, C] -> the activation dtype, on our cast kernel whenever its layout allows (f32, contiguous, C % 4 == 0);
    otherwise a library copy."""
    adt = act_dtype()
    if x.dtype == adt and x.is_contiguous():
        return x
    if x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous() and x.shape[1] % 4 == 0 and x.data_ptr() % 16 == 0:
        return cast_rows(x)
    return x.to(adt).contiguous()


def attention_fwd(qkv, B, T, H, scale):
    o = torch.empty(B * T, H * 64, dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty(B, H, T, dtype=torch.float32, device=qkv.device)
    _lib.call("act_attention_fwd", qkv, B, T, H, 64, float(scale), o, lse, _io32(qkv))
    _count()
    return o, lse


def attention_bwd(qkv, o, do, lse, B, T, H, scale):
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    assert o.dtype == qkv.dtype and do.dtype == qkv.dtype
    _lib.call("act_attention_bwd", qkv, o, do, lse, B, T, H, 64, float(scale), dqkv, delta, _io32(qkv))
    _count(2)
    return dqkv


def colsum(x, out):
    M, N = x.shape
    if x.is_contiguous() and N % 8 == 0 and N <= 2048 and (x.dtype == torch.bfloat16 or M >= 4096):
        _lib.call("act_colsum_bf16_dense", x, _lib.ctypes.c_int64(M), N, out, _io32(x))
        _count()
        return out
    _lib.call("act_colsum", _p(x), int(x.dtype == torch.float32), M, N, x.stride(0), out)
    _count()
    return out


def cosine_loss(student, teacher, eps=1e-8, want_grad=True):
    R, C = student.shape
    loss = torch.empty(1, dtype=torch.float32, device=student.device)
    grad = torch.empty_like(student) if want_grad else None
    _lib.call("act_cosine_loss", student, teacher, R, C, float(eps), loss, grad)
    _count()
    return loss, grad


def zero_(t):
    """cudaMemsetAsync on the current stream (a memset node inside a captured graph, not a fill kernel)."""
    assert t.is_contiguous()
    _lib.call("act_zero", t, _lib.ctypes.c_int64(t.numel() * t.element_size()))
    return t


def accumulate_(dst, src):
    """dst += src (f32, same numel, contiguous)."""
    assert dst.dtype == torch.float32 and src.dtype == torch.float32 and dst.numel() == src.numel()
    _lib.call("act_accumulate", dst, src.contiguous(), _lib.ctypes.c_int64(dst.numel()))
    _count()
    return dst


def scale_by_(x, scalar):
    """x *= scalar, scalar a device f32 tensor with one element."""
    assert x.dtype == torch.float32 and x.is_contiguous() and scalar.numel() == 1
    _lib.call("act_scale_by", x, scalar.float(), _lib.ctypes.c_int64(x.numel()))
    _count()
    return x


_PW_SCRATCH = {}


def pointwise_loss(student, teacher, kind, want_grad=True):
    """mean over all elements of (s - t)^2 (kind 'l2') or SmoothL1(s - t) (kind 'smoothl1') -> (loss f32 [], grad or None)."""
    s, t = _f32c(student), _f32c(teacher)
    assert s.shape == t.shape
    key = str(s.device)
    sc = _PW_SCRATCH.get(key)
    if sc is None:
        sc = (torch.empty(256, dtype=torch.float32, device=s.device), torch.zeros(1, dtype=torch.int32, device=s.device))
        _PW_SCRATCH[key] = sc
    loss = torch.empty((), dtype=torch.float32, device=s.device)
    grad = torch.empty_like(s) if want_grad else None
    _lib.call("act_pointwise_loss", s, t, _lib.ctypes.c_int64(s.numel()), {"l2": 0, "smoothl1": 1}[kind], sc[0], sc[1], loss,
              grad)
    _count()
    return loss, grad


def adamw(param, grad, exp_avg, exp_avg_sq, shadow, n_decay, hyper):
    """grad: the flat fp32 gradient, or its all-reduced bf16 copy (N>1, dp.sync_gradients)."""
    name = "act_adamw_bf16grad" if grad.dtype == torch.bfloat16 else "act_adamw"
    _lib.call(name, _p(param), _p(grad), exp_avg, exp_avg_sq, shadow, _lib.ctypes.c_int64(param.numel()),
              _lib.ctypes.c_int64(n_decay), hyper)
    _count()


def cast_flat_(src, dst):
    """dst (bf16, flat) = bf16(src) (f32, flat; numel % 4 == 0) on our cast kernel."""
    n = src.numel()
    assert dst.numel() == n and n % 4 == 0 and src.dtype == torch.float32 and dst.dtype == torch.bfloat16
    _lib.call("act_cast_rows", src, 1, n, None, 1, _p(dst), 0, None)
    _count()
    return dst



# ------------------------------------------------------------------------------ token plumbing (csrc/tokens.cu)
def pos_mlp1_fwd(x, W, b, out_dtype=torch.bfloat16):
    """GELU(Linear(3,128)(x)): x f32 [R,3] -> [R,128] bf16 (GEMM operand) or f32."""
    x = _f32c(x)
    R = x.shape[0]
    out = torch.empty(R, 128, dtype=out_dtype, device=x.device)
    _lib.call("act_pos_mlp1_fwd", x, W, b, R, _p(out), int(out_dtype == torch.float32))
    _count()
    return out


def pos_mlp1_bwd(da, x, W, b, dW, db):
    """Accumulates dW [128,3], db [128] from da [R,128] (bf16 or f32); the pre-activation is recomputed from x."""
    R = x.shape[0]
    _lib.call("act_pos_mlp1_bwd", _p(da), int(da.dtype == torch.float32), x, W, b, R, dW, db)
    _count()


def mask_block(center, index, num_mask):
    """center f32 [B,G,3], index i32 [B] (device) -> bool [B,G]: the num_mask centres nearest to centre index[b] (act.py:215-243)."""
    center = _f32c(center)
    B, G, _ = center.shape
    assert index.dtype == torch.int32 and index.is_cuda and index.numel() == B
    mask = torch.empty(B, G, dtype=torch.uint8, device=center.device)
    _lib.call("act_mask_block", center, index, B, G, int(num_mask), mask)
    _count()
    return mask.view(torch.bool)


def mask_order(mask):
    """bool / u8 [B,G] -> i64 [B,G]: visible (mask == 0) group indices in original order, then the masked ones."""
    B, G = mask.shape
    m = mask.contiguous()
    if m.dtype == torch.bool:
        m = m.view(torch.uint8)
    order = torch.empty(B, G, dtype=torch.int64, device=mask.device)
    _lib.call("act_mask_order", m, B, G, order)
    _count()
    return order


def permute_groups(nb, center, order, n_vis, want_nb=True):
    """-> (nb_perm [B*G, k, 3] visible rows of all clouds first | None, center_sorted [B,G,3], vis_center [B*n_vis,3])."""
    B, G, _ = center.shape
    center = _f32c(center)
    dev = center.device
    nb_perm = None
    rf = 0
    if want_nb:
        nb = _f32c(nb)
        rf = nb[0, 0].numel()
        nb_perm = torch.empty((B * G,) + tuple(nb.shape[2:]), dtype=torch.float32, device=dev)
    cs = torch.empty(B, G, 3, dtype=torch.float32, device=dev)
    vc = torch.empty(B * n_vis, 3, dtype=torch.float32, device=dev)
    _lib.call("act_permute_groups", nb if want_nb else None, center, order, B, G, rf, n_vis, nb_perm, cs, vc)
    _count()
    return nb_perm, cs, vc


def assemble_rows(src, fill, B, n, T, fill_first, src_T=None, src_off=0):
    """src f32 [B, src_T, C] (rows [src_off, src_off+n) of every cloud) + the parameter row `fill` -> out f32 [B,T,C]."""
    C = fill.numel()
    out = torch.empty(B, T, C, dtype=torch.float32, device=fill.device)
    _lib.call("act_assemble_rows", src, fill, B, n, T, C, int(fill_first), n if src_T is None else src_T, src_off, out)
    _count()
    return out


def assemble_rows_bwd(dout, B, n, T, fill_first, want_dsrc, dfill, src_T=None, src_off=0):
    C = dout.shape[-1]
    src_T = n if src_T is None else src_T
    dsrc = torch.empty(B, src_T, C, dtype=torch.float32, device=dout.device) if want_dsrc else None
    _lib.call("act_assemble_rows_bwd", dout, B, n, T, C, int(fill_first), src_T, src_off, dsrc, dfill)
    _count()
    return dsrc


def gather_rows(src, order, j0, cnt):
    """src f32 [B,G,C] -> [B, cnt, C] = src[b, order[b, j0:j0+cnt]] (order None: the contiguous slice)."""
    B, G, C = src.shape
    out = torch.empty(B, cnt, C, dtype=torch.float32, device=src.device)
    _lib.call("act_gather_rows", _f32c(src), order, B, G, C, int(j0), int(cnt), out)
    _count()
    return out


def embedding_bf16(table, label):
    """table bf16 [V,C], label i32 [R] -> bf16 [R,C]."""
    V, C = table.shape
    R = label.numel()
    out = torch.empty(R, C, dtype=torch.bfloat16, device=table.device)
    _lib.call("act_embedding_bf16", table, label.contiguous(), R, V, C, out)
    _count()
    return out


def drop_path_gates(seed, keep, B, draw_id=0):
    """seed: device int64 [1]; keep f32 [L] -> gates f32 [L,B] = floor(keep + U) / keep."""
    L = keep.numel()
    gates = torch.empty(L, B, dtype=torch.float32, device=keep.device)
    _lib.call("act_drop_path_gates", seed, keep, L, B, int(draw_id), gates)
    _count()
    return gates


# ------------------------------------------------------------------------------ mini-PointNet pieces
def pn_moments(points):
    M = points.shape[0]
    out = torch.empty(9, dtype=torch.float64, device=points.device)
    _lib.call("act_pn_moments", points, _lib.ctypes.c_int64(M), out)
    _count()
    return out


def pn_conv1(points, W, b, relu=True):
    M = points.shape[0]
    out = torch.empty(M, 128, dtype=act_dtype(), device=points.device)
    _lib.call("act_pn_conv1", points, W, b, _lib.ctypes.c_int64(M), int(relu), out, _io32(out))
    _count()
    return out


def group_max(x, k, want_bf16=True, want_f32=False, want_arg=True):
    Mk, C = x.shape
    G = Mk // k
    ob = torch.empty(G, C, dtype=x.dtype, device=x.device) if want_bf16 else None
    of = torch.empty(G, C, dtype=torch.float32, device=x.device) if want_f32 else None
    arg = torch.empty(G, C, dtype=torch.uint8, device=x.device) if want_arg else None
    _lib.call("act_group_max", x, G, k, C, ob, of, arg, _io32(x))
    _count()
    return ob, of, arg


def group_max_bwd(dout, arg, k, out=None):
    G, C = dout.shape
    acc = out is not None
    if out is None:
        out = torch.empty(G * k, C, dtype=act_dtype(), device=dout.device)
    _lib.call("act_group_max_bwd", dout, arg, G, k, C, int(acc), out, _io32(out))
    _count()
    return out


def group_sum(x, k, want_bf16=True, want_f32=False):
    Mk, C = x.shape
    G = Mk // k
    ob = torch.empty(G, C, dtype=x.dtype, device=x.device) if want_bf16 else None
    of = torch.empty(G, C, dtype=torch.float32, device=x.device) if want_f32 else None
    _lib.call("act_group_sum", x, G, k, C, ob, of, _io32(x))
    _count()
    return ob, of


def bn_stats(x):
    M, C = x.shape
    s = torch.empty(2, C, dtype=torch.float32, device=x.device)
    _lib.call("act_bn_stats", x, _lib.ctypes.c_int64(M), C, s[0], s[1], _io32(x))
    _count()
    return s[0], s[1]


def bn_apply(x, scale, shift, relu=True, out=None):
    M, C = x.shape
    if out is None:
        out = torch.empty_like(x)
    _lib.call("act_bn_apply", x, scale, shift, _lib.ctypes.c_int64(M), C, int(relu), out, _io32(x))
    _count()
    return out


def bn_bwd(dz, x, mean, rstd, gamma):
    """-> (dh [M,C], sum_dz (= dbeta), sum_dz_xhat (= dgamma)).  dz may hold only the first M_dz <= M rows: the gradient of
    the remaining rows is zero by construction and is neither stored nor read."""
    M, C = x.shape
    M_dz = dz.shape[0]
    s = torch.empty(2, C, dtype=torch.float32, device=x.device)
    assert dz.dtype == x.dtype and M_dz <= M and dz.is_contiguous()
    _lib.call("act_bn_bwd_stats", dz, x, mean, rstd, _lib.ctypes.c_int64(M_dz), C, s[0], s[1], _io32(x))
    dh = torch.empty_like(x)
    _lib.call("act_bn_bwd_apply", dz, x, mean, rstd, gamma, s[0], s[1], _lib.ctypes.c_int64(M), _lib.ctypes.c_int64(M_dz), C,
              dh, _io32(x))
    _count(2)
    return dh, s[0], s[1]


def relu_mask_(d, a):
    """d *= (a > 0), in place; d, a contiguous, same dtype (activation dtype)."""
    assert d.dtype == a.dtype and d.is_contiguous() and a.is_contiguous() and d.numel() == a.numel()
    _lib.call("act_relu_mask", d, a, _lib.ctypes.c_int64(d.numel()), _io32(d))
    _count()
    return d


def pn_conv1_bwd(dz, points, W, b, mean, rstd, gamma, dW, db):
    """Accumulates dW [128,3], db [128]; returns (dbeta, dgamma) of BatchNorm1."""
    M = points.shape[0]
    s = torch.empty(2, 128, dtype=torch.float32, device=points.device)
    _lib.call("act_pn_conv1_bwd", dz, points, W, b, mean, rstd, gamma, _lib.ctypes.c_int64(M), s[0], s[1], dW, db, _io32(dz))
    _count(2)
    return s[0], s[1]


def pn_bn1_fold(mom9, M, W, b, gamma, beta, eps, momentum, rmean, rvar, nbt):
    """-> (Wf [128,3], bf [128], mean [128], rstd [128]); running stats / counter updated in place."""
    dev = W.device
    o = torch.empty(128 * 6, dtype=torch.float32, device=dev)
    Wf, bf, mean, rstd = o[:384].view(128, 3), o[384:512], o[512:640], o[640:768]
    _lib.call("act_pn_bn1_fold", mom9, _lib.ctypes.c_int64(M), W, b, gamma, beta, float(eps), float(momentum), rmean,
              rvar, nbt, Wf, bf, mean, rstd)
    _count()
    return Wf, bf, mean, rstd


def bn_finalize(s1, s2, M, gamma, beta, eps, momentum, rmean, rvar, nbt):
    """-> (scale, shift, mean, rstd) [C] each; running stats / counter updated in place."""
    C = gamma.numel()
    o = torch.empty(4, C, dtype=torch.float32, device=gamma.device)
    _lib.call("act_bn_finalize", s1, s2, _lib.ctypes.c_int64(M), C, gamma, beta, float(eps), float(momentum), rmean, rvar,
              nbt, o[0], o[1], o[2], o[3])
    _count()
    return o[0], o[1], o[2], o[3]


# ------------------------------------------------------------------------------ frozen teacher (DGCNN) pieces
def dgcnn_edge_gn(pq, idx4, gamma, beta, B, G, Cp, eps, slope, out_view):
    """out_view: bf16 [B*G, Cp] view (row pitch = stride(0)) receiving max_k LeakyReLU(GroupNorm(P[nbr] + Q[self]))."""
    assert pq.dtype == torch.float32 and pq.is_contiguous() and pq.shape == (B * G, 2 * Cp)
    assert out_view.dtype == torch.bfloat16 and out_view.stride(1) == 1
    _lib.call("act_dgcnn_edge_gn", pq, idx4, gamma, beta, B, G, Cp, idx4.shape[-1], 4, float(eps), float(slope),
              _p(out_view), out_view.stride(0))
    _count()
    return out_view


def gn_rows(x, gamma, beta, B, R, eps, slope, noise=None, seed=None):
    """GroupNorm(4)+LeakyReLU over x bf16 [B*R, C].  noise/seed None -> activations f32 [B*R, C];
    noise f32 [B*R, C] -> labels i32 [B*R] = argmax(activation + noise);  seed (device int64 [1]) -> the same with the
    gumbel noise drawn inside the kernel (nothing [B*R, C]-sized is materialised)."""
    C = x.shape[1]
    stats = torch.empty(B, 4, 2, dtype=torch.float32, device=x.device)
    if noise is None and seed is None:
        out = torch.empty(B * R, C, dtype=torch.float32, device=x.device)
        _lib.call("act_gn_rows", x, gamma, beta, B, R, C, 4, float(eps), float(slope), stats, out, None, None, None)
        _count(2)
        return out
    label = torch.empty(B * R, dtype=torch.int32, device=x.device)
    if noise is not None:
        seed = None
    else:
        assert seed.dtype == torch.int64 and seed.is_cuda
    _lib.call("act_gn_rows", x, gamma, beta, B, R, C, 4, float(eps), float(slope), stats, None, noise, seed, label)
    _count(2)
    return label


def vit_ln1_fwd(x, pos_tok, tok, ppos, gamma, beta, eps, B, G, P, keep=None, seed=None, draw_id=0, p_drop=0.0):
    """Entry of a VPT-deep prompted ViT block (include/act_b200.h: act_vit_ln1_fwd) ->
    (xs f32 [B*G,C] token residual stream, h_tok bf16 [B*G,C], h_prm bf16 [B*P,C])."""
    C = pos_tok.shape[-1]
    dev = pos_tok.device
    xs = torch.empty(B * G, C, dtype=torch.float32, device=dev)
    h_tok = torch.empty(B * G, C, dtype=torch.bfloat16, device=dev)
    h_prm = torch.empty(B * P, C, dtype=torch.bfloat16, device=dev)
    _lib.call("act_vit_ln1_fwd", x, pos_tok, tok, ppos, keep, seed, draw_id, float(p_drop), gamma, beta, float(eps), B, G,
              P, C, xs, _p(h_tok), _p(h_prm))
    _count()
    return xs, h_tok, h_prm


def attention_prefix_fwd(qkv_t, kv_p, B, G, P, H, scale):
    """Queries = the G token rows of qkv_t [B*G, 3*H*64]; keys/values = P prompt rows of kv_p [B*P, 2*H*64] + the
    token rows -> o bf16 [B*G, H*64]."""
    o = torch.empty(B * G, H * 64, dtype=torch.bfloat16, device=qkv_t.device)
    _lib.call("act_attention_prefix_fwd", qkv_t, kv_p, B, G, P, H, 64, float(scale), o)
    _count()
    return o


# ------------------------------------------------------------------ Stage-I dVAE: trainable DGCNN layers (csrc/dgcnn_train.cu)
def dgcnn_edge_gn_train_fwd(pq, idx4, gamma, beta, B, G, Cp, eps, slope):
    """max_k LeakyReLU(GroupNorm(P[nbr] + Q[self])) -> (out f32 [B*G, Cp], argj u8 [B*G, Cp], stats f32 [B,4,2])."""
    assert pq.dtype == torch.float32 and pq.is_contiguous() and pq.shape == (B * G, 2 * Cp)
    out = torch.empty(B * G, Cp, dtype=torch.float32, device=pq.device)
    argj = torch.empty(B * G, Cp, dtype=torch.uint8, device=pq.device)
    stats = torch.empty(B, 4, 2, dtype=torch.float32, device=pq.device)
    _lib.call("act_dgcnn_edge_gn_train_fwd", pq, idx4, _f32c(gamma), _f32c(beta), B, G, Cp, idx4.shape[-1], 4, float(eps),
              float(slope), out, Cp, argj, stats)
    _count()
    return out, argj, stats


def dgcnn_edge_gn_train_bwd(pq, idx4, argj, stats, gamma, beta, dout, B, G, Cp, slope, dgamma, dbeta):
    """-> dpq f32 [B*G, 2*Cp]; dgamma / dbeta (f32 [Cp]) are accumulated into."""
    dout = _f32c(dout)
    dpq = torch.empty(B * G, 2 * Cp, dtype=torch.float32, device=pq.device)
    sums = torch.empty(B, 4, 2, dtype=torch.float32, device=pq.device)
    _lib.call("act_dgcnn_edge_gn_train_bwd", pq, idx4, argj, stats, _f32c(gamma), _f32c(beta), dout, Cp, B, G, Cp,
              idx4.shape[-1], 4, float(slope), sums, dpq, dgamma, dbeta)
    _count(2)
    return dpq


def gn_rows_train_fwd(x, gamma, beta, B, R, eps, slope):
    """LeakyReLU(GroupNorm(4)(x)) for x f32 [B*R, C] -> (out f32, stats f32 [B,4,2])."""
    x = _f32c(x)
    C = x.shape[1]
    out = torch.empty_like(x)
    stats = torch.empty(B, 4, 2, dtype=torch.float32, device=x.device)
    _lib.call("act_gn_rows_train_fwd", x, _f32c(gamma), _f32c(beta), B, R, C, 4, float(eps), float(slope), stats, out)
    _count(2)
    return out, stats


def gn_rows_train_bwd(x, stats, gamma, beta, dy, B, R, slope, dgamma, dbeta):
    dy = _f32c(dy)
    C = x.shape[1]
    dx = torch.empty_like(x)
    sums = torch.empty(B, 4, 2, dtype=torch.float32, device=x.device)
    _lib.call("act_gn_rows_train_bwd", x, stats, _f32c(gamma), _f32c(beta), dy, B, R, C, 4, float(slope), sums, dx,
              dgamma, dbeta)
    _count(2)
    return dx


# ------------------------------------------------------------ Stage-I gumbel-softmax + KL (csrc/gumbel.cu)
GUMBEL_V = (1024, 2048, 4096, 8192, 16384)


def _tau_args(tau):
    """tau: python float, or a 0-dim / 1-element f32 device tensor (the engine's staged schedule) -> (ptr, value)."""
    if isinstance(tau, torch.Tensor):
        assert tau.is_cuda and tau.dtype == torch.float32 and tau.numel() == 1
        return tau, 0.0
    return None, float(tau)


def gumbel_softmax_fwd(logits, tau, noise=None, seed=None, draw_id=0):
    """logits f32 [R,V] -> (y [R,V] in the activation dtype = softmax((logits + gumbel) / tau), lse f32 [R])."""
    R, V = logits.shape
    y = torch.empty(R, V, dtype=act_dtype(), device=logits.device)
    lse = torch.empty(R, dtype=torch.float32, device=logits.device)
    tp, tv = _tau_args(tau)
    _lib.call("act_gumbel_softmax_fwd", logits, noise, seed, int(draw_id), tp, tv, R, V, int(y.dtype == torch.bfloat16),
              _p(y), lse)
    _count()
    return y, lse


def softmax_colmean(logits, lse, B, G):
    V = logits.shape[1]
    qbar = torch.empty(B, V, dtype=torch.float32, device=logits.device)
    _lib.call("act_softmax_colmean", logits, lse, B, G, V, qbar)
    _count()
    return qbar


_KL_SCRATCH = {}


def kl_uniform_fwd(qbar):
    B, V = qbar.shape
    key = (str(qbar.device), B)
    sc = _KL_SCRATCH.get(key)
    if sc is None:       # first use happens in the engine's eager warm-up, outside any capture
        sc = (torch.empty(B, dtype=torch.float32, device=qbar.device), torch.zeros(1, dtype=torch.int32, device=qbar.device))
        _KL_SCRATCH[key] = sc
    loss = torch.empty((), dtype=torch.float32, device=qbar.device)
    _lib.call("act_kl_uniform_fwd", qbar, B, V, sc[0], sc[1], loss)
    _count()
    return loss


def kl_uniform_bwd(qbar, gout):
    B, V = qbar.shape
    dq = torch.empty_like(qbar)
    _lib.call("act_kl_uniform_bwd", qbar, _f32c(gout), B, V, dq)
    _count()
    return dq


def gumbel_softmax_bwd(logits, lse, y, dy, tau, dqbar, G):
    R, V = logits.shape
    dl = torch.empty_like(logits)
    tp, tv = _tau_args(tau)
    if dy is not None:
        dy = dy.contiguous()
        if dy.dtype != y.dtype:
            dy = dy.to(y.dtype)
    _lib.call("act_gumbel_softmax_bwd", logits, lse, _p(y) if dy is not None else None, _p(dy) if dy is not None else None,
              int(y.dtype == torch.bfloat16), tp, tv if dy is not None else 1.0, dqbar, R, G, V, dl)
    _count()
    return dl


def edge_weight_fwd(W):
    """W f32 [Cp, 2*Cin] -> [2*Cp, Cin] = [Wa ; Wb - Wa] in the activation dtype (the GEMM's B operand)."""
    Cp, Cin = W.shape[0], W.shape[1] // 2
    out = torch.empty(2 * Cp, Cin, dtype=act_dtype(), device=W.device)
    _lib.call("act_edge_weight_fwd", W, Cp, Cin, int(out.dtype == torch.bfloat16), _p(out))
    _count()
    return out


def edge_weight_bwd(dWp, dW):
    Cp, Cin = dW.shape[0], dW.shape[1] // 2
    assert dWp.shape == (2 * Cp, Cin) and dWp.dtype == torch.float32 and dW.is_contiguous()
    _lib.call("act_edge_weight_bwd", dWp, Cp, Cin, dW)
    _count()


def fold_input_fwd(z_g, coarse, weight, c_g, seed):
    """FoldingNet final_conv.0 as a broadcast sum (csrc/folding.cu): z_g f32 [BG,C], coarse f32 [BG,M,3], weight f32
    [C, c_g + 5] (the full conv weight; only its last 5 columns are read), seed f32 [S,2] -> z [BG*M*S, C] (act dtype)."""
    BG, C = z_g.shape
    M, S = coarse.shape[1], seed.shape[0]
    z = torch.empty(BG * M * S, C, dtype=act_dtype(), device=z_g.device)
    _lib.call("act_fold_input_fwd", z_g, coarse, _vp(weight.data_ptr() + 4 * c_g), weight.stride(0), seed, BG, M, S, C,
              int(z.dtype == torch.bfloat16), _p(z))
    _count()
    return z


def fold_input_bwd(dz, coarse, weight, c_g, seed, dweight):
    """-> (dz_g f32 [BG,C], dcoarse f32 [BG,M,3]); the last 5 columns of dweight f32 [C, c_g + 5] are accumulated into."""
    BG, M = coarse.shape[:2]
    C, S = weight.shape[0], seed.shape[0]
    dz_g = torch.empty(BG, C, dtype=torch.float32, device=dz.device)
    dcoarse = torch.empty(BG, M, 3, dtype=torch.float32, device=dz.device)
    assert dweight.stride(0) == weight.stride(0) and dweight.dtype == torch.float32
    _lib.call("act_fold_input_bwd", _p(dz), int(dz.dtype == torch.bfloat16), coarse, _vp(weight.data_ptr() + 4 * c_g),
              weight.stride(0), seed, BG, M, S, C, dz_g, dcoarse, _vp(dweight.data_ptr() + 4 * c_g))
    _count()
    return dz_g, dcoarse


# ------------------------------------------------------------------------------------- input augmentation
def scale_translate_(pc, st):
    """In place: pc[b,n,:] = pc[b,n,:] * st[b,:3] + st[b,3:]  (data_transforms.py:20-34); pc f32 [B,N,3], st f32 [B,6]."""
    assert pc.dtype == torch.float32 and pc.is_contiguous() and pc.dim() == 3 and pc.shape[2] == 3
    assert st.dtype == torch.float32 and st.is_contiguous() and st.shape == (pc.shape[0], 6)
    _lib.call("act_scale_translate", pc, st, pc.shape[0], pc.shape[1])
    _count()
    return pc


def subsample_norm(raw, sel):
    """ShapeNet.__getitem__'s random_sample + pc_norm for a batch: raw f32 [B,Nraw,3], sel i32 [B,num] -> f32 [B,num,3]."""
    raw = _f32c(raw)
    B, Nraw, _ = raw.shape
    num = sel.shape[1]
    out = torch.empty(B, num, 3, dtype=torch.float32, device=raw.device)
    _lib.call("act_subsample_norm", raw, sel.to(torch.int32).contiguous(), B, Nraw, num, out)
    _count()
    return out


def mask_rand(seed, B, G, num_mask):
    """Device-side random mask: bool [B,G] with exactly num_mask ones per row; seed = device int64 [1]."""
    mask = torch.empty(B, G, dtype=torch.uint8, device=seed.device)
    _lib.call("act_mask_rand", seed, B, G, int(num_mask), mask)
    _count()
    return mask.view(torch.bool)
