"""sys.modules stand-ins that let the UNMODIFIED reference import on a CPU-only box.
TEST INFRASTRUCTURE ONLY -- used by oracle/make_golden.py and tests/ (only when /root/reference
exists, i.e. in the authoring container).  Never imported by act_b200/.

The reference (RunpeiDong/ACT) imports nine packages that are not installed here
(SURVEY.md section 8c): easydict, mmcv, termcolor, matplotlib, timm, lightly, knn_cuda,
pointnet2_ops and its own compiled `chamfer` extension.  The three native ones are backed by the
C oracle (oracle/cpu_ref.c); the rest are minimal functional stand-ins of the few symbols the
hot-path files touch.
"""
import math
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cpu_ref

REFERENCE_ROOT = "/root/reference"


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


class _EasyDict(dict):
    """easydict.EasyDict: attribute access, recursive on nested dicts."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(_EasyDict(x) if isinstance(x, dict) else x for x in v)
        dict.__setitem__(self, k, v)
        object.__setattr__(self, k, v)

    __setitem__ = __setattr__


class _DropPath(nn.Module):
    """timm 0.5.4 DropPath (per-sample stochastic depth)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        r = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        r.floor_()
        return x.div(keep) * r


def _trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)


class _NegCos(nn.Module):
    """lightly 1.2.28 NegativeCosineSimilarity."""

    def __init__(self, dim=1, eps=1e-8):
        super().__init__()
        self.dim, self.eps = dim, eps

    def forward(self, x0, x1):
        return -F.cosine_similarity(x0, x1, self.dim, self.eps).mean()


class _KNN(nn.Module):
    """knn_cuda.KNN v0.2 front-end backed by oracle_knn."""

    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k, self._t = k, transpose_mode

    def forward(self, ref, query):
        with torch.no_grad():
            if not self._t:
                ref, query = ref.transpose(1, 2), query.transpose(1, 2)
            d, i = cpu_ref.knn(ref.contiguous().float().numpy(), query.contiguous().float().numpy(), self.k)
            d, i = torch.from_numpy(d), torch.from_numpy(i)
            if not self._t:
                d, i = d.transpose(1, 2).contiguous(), i.transpose(1, 2).contiguous()
        return d, i


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return torch.from_numpy(cpu_ref.gather(features.detach().numpy(), idx.numpy()))

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return torch.from_numpy(cpu_ref.gather_grad(g.contiguous().numpy(), idx.numpy(), ctx.n)), None


def _fps(xyz, npoint):
    return torch.from_numpy(cpu_ref.fps(xyz.detach().contiguous().numpy(), npoint))


def _chamfer_forward(xyz1, xyz2):
    return [torch.from_numpy(a) for a in cpu_ref.chamfer_forward(xyz1.detach().numpy(), xyz2.detach().numpy())]


def _chamfer_backward(xyz1, xyz2, idx1, idx2, g1, g2):
    return [torch.from_numpy(a) for a in cpu_ref.chamfer_backward(
        xyz1.detach().numpy(), xyz2.detach().numpy(), idx1.numpy(), idx2.numpy(),
        g1.contiguous().numpy(), g2.contiguous().numpy())]


_installed = False


def install(native="oracle", reference_root=None):
    """Install the stand-ins and put the reference on sys.path.  Idempotent.
    native = "oracle": knn_cuda / pointnet2_ops / chamfer are backed by the C oracle (CPU);
    native = "dropin": those three import names are NOT stubbed -- the repo's dropin/ directory is put on sys.path instead,
    so the unmodified reference runs on the act_b200 kernels (GPU; tests/test_gpu_dropin_reference.py)."""
    global _installed, REFERENCE_ROOT
    if _installed:
        return
    _installed = True
    if reference_root:
        REFERENCE_ROOT = reference_root
    _mod("easydict").EasyDict = _EasyDict
    mmcv = _mod("mmcv")
    mmcv.utils = _mod("mmcv.utils")
    mmcv.utils.collect_env = lambda: {}
    _mod("termcolor").colored = lambda s, *a, **k: s
    mpl = _mod("matplotlib")
    mpl.pyplot = _mod("matplotlib.pyplot")
    mt = _mod("mpl_toolkits")
    mt.mplot3d = _mod("mpl_toolkits.mplot3d")
    mt.mplot3d.Axes3D = object
    timm = _mod("timm")
    timm.models = _mod("timm.models")
    timm.models.layers = _mod("timm.models.layers")
    timm.models.layers.DropPath = _DropPath
    timm.models.layers.trunc_normal_ = _trunc_normal_
    timm.scheduler = _mod("timm.scheduler")
    timm.scheduler.CosineLRScheduler = object
    def _create_model(name, pretrained=False, **k):
        # stand-in for timm.create_model('vit_base_patch16_384'): random-weight ViT-B blocks (no pretrained weights
        # offline); the reference only touches .blocks / .norm / .embed_dim (models/dvae.py:405-411)
        from .ref_teacher import FakeTimmViT
        return FakeTimmViT(768, 12, 12)
    timm.create_model = _create_model
    lightly = _mod("lightly")
    lightly.loss = _mod("lightly.loss")
    lightly.loss.NegativeCosineSimilarity = _NegCos
    if native == "dropin":
        import os
        dropin = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin")
        if dropin not in sys.path:
            sys.path.insert(0, dropin)
    else:
        _mod("knn_cuda").KNN = _KNN
        p2 = _mod("pointnet2_ops")
        p2.pointnet2_utils = _mod("pointnet2_ops.pointnet2_utils")
        p2.pointnet2_utils.furthest_point_sample = _fps
        p2.pointnet2_utils.gather_operation = _Gather.apply
        ch = _mod("chamfer")
        ch.forward, ch.backward = _chamfer_forward, _chamfer_backward
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # act.py:1243 and dvae.py:300 call .cuda() unconditionally; neutralise on a CPU-only box.
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self


def easydict(d):
    return _EasyDict(d)
