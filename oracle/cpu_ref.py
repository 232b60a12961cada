"""ctypes front-end of oracle/cpu_ref.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (act_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "cpu_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def fps(xyz, m):
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    idx = np.zeros((B, m), dtype=np.int32)
    lib().oracle_fps(_p(xyz), B, N, m, _p(idx))
    return idx


def gather(feat, idx):
    feat = _f32(feat)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, C, N = feat.shape
    M = idx.shape[1]
    out = np.zeros((B, C, M), dtype=np.float32)
    lib().oracle_gather(_p(feat), _p(idx), B, C, N, M, _p(out))
    return out


def gather_grad(gout, idx, N):
    gout = _f32(gout)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, C, M = gout.shape
    out = np.zeros((B, C, N), dtype=np.float32)
    lib().oracle_gather_grad(_p(gout), _p(idx), B, C, N, M, _p(out))
    return out


def knn(ref, query, k):
    """ref [B,N,3], query [B,Q,3] -> (dist [B,Q,k] f32 euclidean ascending, idx [B,Q,k] i64)."""
    ref, query = _f32(ref), _f32(query)
    B, N, _ = ref.shape
    Q = query.shape[1]
    dist = np.zeros((B, Q, k), dtype=np.float32)
    idx = np.zeros((B, Q, k), dtype=np.int64)
    lib().oracle_knn(_p(ref), _p(query), B, N, Q, k, _p(dist), _p(idx))
    return dist, idx


def group(xyz, G, K):
    """Group.forward: -> (neighborhood [B,G,K,3], center [B,G,3], idx [B,G,K] i64, fps_idx [B,G] i32)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    fps_idx = np.zeros((B, G), dtype=np.int32)
    center = np.zeros((B, G, 3), dtype=np.float32)
    idx = np.zeros((B, G, K), dtype=np.int64)
    nb = np.zeros((B, G, K, 3), dtype=np.float32)
    lib().oracle_group(_p(xyz), B, N, G, K, _p(fps_idx), _p(center), _p(idx), _p(nb))
    return nb, center, idx, fps_idx


def chamfer_forward(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.zeros((B, n), np.float32)
    d2 = np.zeros((B, m), np.float32)
    i1 = np.zeros((B, n), np.int32)
    i2 = np.zeros((B, m), np.int32)
    lib().oracle_chamfer_forward(_p(xyz1), _p(xyz2), B, n, m, _p(d1), _p(d2), _p(i1), _p(i2))
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, idx1, idx2, g1, g2):
    xyz1, xyz2, g1, g2 = _f32(xyz1), _f32(xyz2), _f32(g1), _f32(g2)
    idx1 = np.ascontiguousarray(idx1, dtype=np.int32)
    idx2 = np.ascontiguousarray(idx2, dtype=np.int32)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = np.zeros((B, n, 3), np.float32)
    gx2 = np.zeros((B, m, 3), np.float32)
    lib().oracle_chamfer_backward(_p(xyz1), _p(xyz2), _p(idx1), _p(idx2), _p(g1), _p(g2), B, n, m, _p(gx1), _p(gx2))
    return gx1, gx2
