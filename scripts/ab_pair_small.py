"""Dispatch sweep for the K-major forward GEMMs of the transformer stacks: default tile choice vs the CTA-pair kernel on
256 x 256 tiles (persistent=2) and on 256 x 384 tiles (persistent=3), graph-timed: python scripts/ab_pair_small.py"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
from scripts.kbench_graph import graph_time, bf

def sweep(tag, cases):
    for name, fn in cases.items():
        row = {}
        for label, kw in (("auto", {}), ("pair256", dict(persistent=2)), ("pair384", dict(persistent=3))):
            try:
                row[label] = round(graph_time(lambda: fn(**kw)), 2)
            except Exception as e:
                row[label] = str(e)[:30]
        print(tag, name, json.dumps(row), flush=True)

for tag, M, D in (("enc", 3456, 384), ("dec", 8192, 384), ("vit", 8192, 768)):
    x, xh = bf(M, D), bf(M, 4 * D)
    wqkv, wproj, w1, w2 = bf(3 * D, D, scale=.05), bf(D, D, scale=.05), bf(4 * D, D, scale=.05), bf(D, 4 * D, scale=.05)
    bD, b4D = torch.randn(D, device="cuda"), torch.randn(4 * D, device="cuda")
    xs = torch.randn(M, D, device="cuda")
    u = torch.empty(M, 4 * D, dtype=torch.bfloat16, device="cuda")
    a = torch.empty(M, 4 * D, dtype=torch.bfloat16, device="cuda")
    o_qkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device="cuda")
    o_f32 = torch.empty(M, D, device="cuda")
    cases = {
        "qkv_fwd": lambda **k: ops.gemm(x, wqkv, out=o_qkv, **k),
        "proj_fwd": lambda **k: ops.gemm(x, wproj, bias=bD, resid=xs, out=o_f32, **k),
        "fc1_fwd_gelu": (lambda **k: ops.gemm(x, w1, bias=b4D, act=1, preact_out=u, out=a, **k)) if tag != "vit" else
                        (lambda **k: ops.gemm(x, w1, bias=b4D, act=1, out=a, **k)),
        "fc2_fwd": lambda **k: ops.gemm(xh, w2, bias=bD, resid=xs, out=o_f32, **k),
    }
    sweep(tag, cases)
