"""ctypes loader for libact_b200.so (the C ABI declared in include/act_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, this raises.
Tensors cross the boundary as raw device pointers + sizes; the current torch CUDA stream is passed to
every call, so launches are ordered with PyTorch work and are CUDA-graph capturable.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACT_B200_LIB") or os.path.join(_HERE, "libact_b200.so")   # override: A/B runs of two builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "act_b200.h")

_lib = None


class ActB200Error(RuntimeError):
    pass


def build(verbose=False):
    """Compile act_b200/csrc/*.cu for sm_100a into act_b200/libact_b200.so (nvcc cross-compiles on CPU)."""
    import subprocess
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise ActB200Error("building libact_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


def declared_symbols():
    """Names of every function include/act_b200.h declares."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(act_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ActB200Error(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU / PyTorch fallback for the act_b200 kernels)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.act_error_string.restype = ctypes.c_char_p
        _lib.act_error_string.argtypes = [ctypes.c_int]
        _lib.act_version.restype = ctypes.c_int
    return _lib


def check(rc, what):
    if rc != 0:
        raise ActB200Error(f"{what} failed ({rc}): {lib().act_error_string(rc).decode()}")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise ActB200Error("act_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise ActB200Error("act_b200 kernels need contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    """Invoke `name` with the current stream appended; raise on a non-zero return code."""
    fn = getattr(lib(), name)
    conv = []
    for a in args:
        if isinstance(a, torch.Tensor) or a is None:
            conv.append(ptr(a))
        elif isinstance(a, float):
            conv.append(ctypes.c_float(a))
        elif isinstance(a, (ctypes.c_void_p, ctypes.c_float, ctypes.c_double, ctypes.c_int64)):
            conv.append(a)
        else:
            conv.append(ctypes.c_int(int(a)))
    check(fn(*conv, stream()), name)
