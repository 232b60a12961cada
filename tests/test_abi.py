"""CPU: the C-ABI library builds/loads and exports every symbol include/act_b200.h declares (no compute)."""
import os

from act_b200 import _lib


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = _lib.lib()
    syms = _lib.declared_symbols()
    assert "act_fps" in syms and "act_knn" in syms and "act_chamfer_forward" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/act_b200.h but not exported"
    assert lib.act_version() >= 100
    assert b"invalid" in lib.act_error_string(-1)


def test_rejects_bad_arguments_without_gpu():
    import ctypes
    lib = _lib.lib()
    null = ctypes.c_void_p(0)
    assert lib.act_fps(null, 1, 16, 4, null, null, null) == -1
    assert lib.act_knn(null, null, 1, 16, 4, 4, null, null, null, null) == -1
    assert lib.act_chamfer_forward(null, null, 1, 4, 4, null, null, null, null, null) == -1


def test_product_path_has_no_cpu_fallback():
    import pytest
    import torch
    from act_b200 import ops
    with pytest.raises(_lib.ActB200Error):
        ops.furthest_point_sample(torch.zeros(1, 8, 3), 2)     # CPU tensor -> loud failure


def test_product_never_imports_oracle():
    root = os.path.dirname(_lib._HERE)
    for d in ("act_b200", "dropin"):
        for dp, _, fs in os.walk(os.path.join(root, d)):
            for f in fs:
                if f.endswith(".py"):
                    src = open(os.path.join(dp, f)).read()
                    assert "oracle" not in src.replace("no oracle", ""), os.path.join(dp, f)
