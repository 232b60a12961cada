"""act_b200 -- B200 (sm_100a) kernels and host-side modules for ACT's masked-point-modeling hot path.

Layout: csrc/ (CUDA kernels + the C ABI of include/act_b200.h), _lib.py (ctypes loader, no fallback),
ops.py (tensor-level wrappers / autograd Functions), modules.py + models.py (the reference's nn.Module
surface: Group, Encoder, Block, ..., ACT_PointDistillation).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
