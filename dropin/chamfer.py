"""Drop-in for the reference's compiled `chamfer` extension (extensions/chamfer_dist/setup.py:11-19,
chamfer_cuda.cpp:36-39): forward(xyz1, xyz2) -> [dist1, dist2, idx1, idx2];
backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2) -> [grad_xyz1, grad_xyz2]."""
from act_b200.ops import chamfer_backward as backward  # noqa: F401
from act_b200.ops import chamfer_forward as forward    # noqa: F401
