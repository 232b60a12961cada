"""The FMA-contraction order the oracle (and the CUDA kernels) assume for the three tokenizer distance expressions,
pinned against what nvcc actually emits for the upstream SOURCE forms (oracle/fma_probe.cu).

pointnet2_ops and KNN_CUDA are not vendored and cannot run offline, so the bit-exactness of FPS / kNN indices rests on
restating their arithmetic.  The one compiler-dependent part of that arithmetic is which products nvcc (-fmad=true, the
default those packages are built with) fuses into FMAs, and in which order: this test compiles the source forms with
the local nvcc, rebuilds every stored value's expression tree from the PTX, and compares it with the order written out
in oracle/cpu_ref.c (fmaf) and act_b200/csrc/{fps,knn,chamfer}.cu (__fmaf_rn).  CPU-only: nvcc -ptx needs no GPU."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _ptx(tmp_path):
    out = tmp_path / "fma_probe.ptx"
    subprocess.check_call([NVCC, "-ptx", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out),
                           os.path.join(ROOT, "oracle", "fma_probe.cu")])
    return out.read_text()


def _stored_expressions(ptx, entry):
    """Symbolically execute the straight-line body of `entry`: -> {(param index, float offset): expression string}."""
    body = ptx[ptx.index(f".entry {entry}("):]
    body = body[:body.index("ret;")]
    sym, stores = {}, {}
    for line in body.splitlines():
        line = line.strip().rstrip(";")
        m = re.match(r"ld\.param\.u64\s+(%rd\d+), \[\w+_param_(\d+)\]", line)
        if m:
            sym[m.group(1)] = int(m.group(2))
            continue
        m = re.match(r"cvta\.to\.global\.u64\s+(%rd\d+), (%rd\d+)", line)
        if m:
            sym[m.group(1)] = sym[m.group(2)]
            continue
        m = re.match(r"ld\.global(?:\.nc)?\.f32\s+(%f\d+), \[(%rd\d+)(?:\+(\d+))?\]", line)
        if m:
            sym[m.group(1)] = f"p{sym[m.group(2)]}[{int(m.group(3) or 0) // 4}]"
            continue
        m = re.match(r"(sub|mul|add)(?:\.rn)?\.f32\s+(%f\d+), (\S+), (\S+)$", line)
        if m:
            a, b = (sym.get(x, x) for x in (m.group(3), m.group(4)))
            sym[m.group(2)] = f"{m.group(1)}({a},{b})"
            continue
        m = re.match(r"fma\.rn\.f32\s+(%f\d+), (\S+), (\S+), (\S+)$", line)
        if m:
            a, b, c = (sym.get(x, x) for x in (m.group(2), m.group(3), m.group(4)))
            sym[m.group(1)] = f"fma({a},{b},{c})"
            continue
        m = re.match(r"st\.global\.f32\s+\[(%rd\d+)(?:\+(\d+))?\], (%f\d+)", line)
        if m:
            stores[(sym[m.group(1)], int(m.group(2) or 0) // 4)] = sym[m.group(3)]
    return stores


@pytest.fixture(scope="module")
def ptx(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    return _ptx(tmp_path_factory.mktemp("fma"))


def test_fps_expression_is_mul_y_then_fma_x_then_fma_z(ptx):
    """oracle/cpu_ref.c oracle_fps: mag = fmaf(z,z, fmaf(x,x, y*y)); d = fmaf(dz,dz, fmaf(dx,dx, dy*dy))."""
    st = _stored_expressions(ptx, "probe_fps")
    x2, y2, z2 = "p1[0]", "p1[1]", "p1[2]"
    assert st[(2, 0)] == f"fma({z2},{z2},fma({x2},{x2},mul({y2},{y2})))"
    dx, dy, dz = (f"sub(p1[{i}],p0[{i}])" for i in range(3))
    assert st[(2, 1)] == f"fma({dz},{dz},fma({dx},{dx},mul({dy},{dy})))"


def test_knn_expression_accumulates_x_y_z_with_fma(ptx):
    """oracle/cpu_ref.c oracle_knn: d = fmaf(dz,dz, fmaf(dy,dy, dx*dx)) -- nvcc emits fma(dx,dx,+0), which rounds exactly
    like the plain product."""
    st = _stored_expressions(ptx, "probe_knn")
    dx, dy, dz = (f"sub(p0[{i}],p1[{i}])" for i in range(3))
    assert st[(2, 0)] == f"fma({dz},{dz},fma({dy},{dy},fma({dx},{dx},0f00000000)))"


def test_chamfer_expression_matches_fps_order(ptx):
    """oracle/cpu_ref.c oracle_chamfer_forward (reference extensions/chamfer_dist/chamfer.cu:42-46)."""
    st = _stored_expressions(ptx, "probe_chamfer")
    dx, dy, dz = (f"sub(p1[{i}],p0[{i}])" for i in range(3))
    assert st[(2, 0)] == f"fma({dz},{dz},fma({dx},{dx},mul({dy},{dy})))"


def test_oracle_source_states_the_same_order():
    """The order the probe pins is the one written in oracle/cpu_ref.c (guards against the two drifting apart)."""
    src = open(os.path.join(ROOT, "oracle", "cpu_ref.c")).read()
    assert "fmaf(dz, dz, fmaf(dx, dx, dy * dy))" in src            # FPS / Chamfer form
    assert re.search(r"fmaf\(dz, dz, fmaf\(dy, dy, dx \* dx\)\)", src)   # kNN form
