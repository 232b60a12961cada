"""Drop-in for pointnet2_ops.pointnet2_utils (erikwijmans/Pointnet2_PyTorch) -- only the two entry points
the reference calls (utils/misc.py:44-45, tools/runner_finetune.py:155-157), on act_b200 kernels."""
from act_b200 import ops as _ops


def furthest_point_sample(xyz, npoint):
    """xyz [B,N,3] CUDA f32 contiguous -> int32 [B,npoint] (non-differentiable)."""
    return _ops.furthest_point_sample(xyz, npoint)


def gather_operation(features, idx):
    """features [B,C,N] f32, idx [B,M] i32 -> [B,C,M]; differentiable w.r.t. features."""
    return _ops.gather_operation(features, idx)
