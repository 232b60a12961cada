"""Tile-shape / scheduling sweep for the small transformer GEMMs (graph-timed)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
from scripts.kbench_graph import graph_time, bf

for tag, M in (("enc", 3456), ("dec", 8192)):
    x384, x1536 = bf(M, 384), bf(M, 1536)
    wqkv, wproj, w1, w2 = bf(1152, 384, scale=.05), bf(384, 384, scale=.05), bf(1536, 384, scale=.05), bf(384, 1536, scale=.05)
    b384, b1536 = torch.randn(384, device="cuda"), torch.randn(1536, device="cuda")
    xs = torch.randn(M, 384, device="cuda")
    u = torch.empty(M, 1536, dtype=torch.bfloat16, device="cuda")
    a = torch.empty(M, 1536, dtype=torch.bfloat16, device="cuda")
    o_qkv = torch.empty(M, 1152, dtype=torch.bfloat16, device="cuda")
    o_f32 = torch.empty(M, 384, device="cuda")
    o384 = torch.empty(M, 384, dtype=torch.bfloat16, device="cuda")
    cases = {
        "qkv_fwd": lambda **k: ops.gemm(x384, wqkv, out=o_qkv, **k),
        "proj_fwd": lambda **k: ops.gemm(x384, wproj, bias=b384, resid=xs, out=o_f32, **k),
        "fc1_fwd_gelu": lambda **k: ops.gemm(x384, w1, bias=b1536, act=1, preact_out=u, out=a, **k),
        "fc2_fwd": lambda **k: ops.gemm(x1536, w2, bias=b384, resid=xs, out=o_f32, **k),
        "fc2_dgrad_gelugrad": lambda **k: ops.gemm(x384, w2, b_mn=True, mul_in=u, mul_mode=1, out=a, **k),
        "fc1_dgrad": lambda **k: ops.gemm(x1536, w1, b_mn=True, out=o384, **k),
    }
    for name, fn in cases.items():
        row = {}
        for label, kw in (("v1_bn128", dict(persistent=0, block_n=128)), ("v1_bn64", dict(persistent=0, block_n=64)),
                          ("p_bn128", dict(persistent=1, block_n=128)), ("p_bn256", dict(persistent=1, block_n=256))):
            try:
                row[label] = round(graph_time(lambda: fn(**kw)), 2)
            except Exception as e:
                row[label] = str(e)[:40]
        print(tag, name, json.dumps(row), flush=True)
