"""Tensor-level wrappers over the C ABI (include/act_b200.h).  Every function here launches hand-written
sm_100a kernels from libact_b200.so on the current CUDA stream; none has a PyTorch/CPU fallback."""
import torch

from . import _lib

LAUNCHES = 0  # number of libact_b200 kernel launches issued through this module (bench.py reads it)


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


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.bfloat16, bias=None, act=ACT_NONE,
         preact_out=None, mul_in=None, mul_mode=MUL_NONE, resid=None, alpha=1.0, splits=1, block_n=0):
    """out[M,N] = epilogue(alpha * A . B^T) on the tcgen05 GEMM (include/act_b200.h: act_gemm_bf16).
    a: bf16 [M,K] (or [K,M] if a_mn);  b: bf16 [N,K] (or [K,N] if b_mn).  2-D, last-dim contiguous."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    K, M = (a.shape if a_mn else a.shape[::-1])
    Kb, N = (b.shape if b_mn else b.shape[::-1])
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    if out is None:
        out = (torch.zeros if splits > 1 else torch.empty)(M, N, dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in (torch.bfloat16, torch.float32)
    if preact_out is not None:
        assert preact_out.dtype == torch.bfloat16 and preact_out.stride(0) == out.stride(0)
    _lib.call("act_gemm_bf16", _lib.ctypes.c_void_p(a.data_ptr()), _lib.ctypes.c_void_p(b.data_ptr()), M, N, K,
              int(a_mn), int(b_mn), a.stride(0), b.stride(0), _lib.ctypes.c_void_p(out.data_ptr()), out.stride(0),
              int(out.dtype == torch.float32), bias, int(act), preact_out, mul_in,
              mul_in.stride(0) if mul_in is not None else 0, int(mul_mode),
              _lib.ctypes.c_void_p(resid.data_ptr()) if resid is not None else None,
              resid.stride(0) if resid is not None else 0, float(alpha), int(splits), int(block_n))
    _count()
    return out
