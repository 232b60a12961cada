"""A/B of the one-tile GEMM kernel with 4 vs 8 epilogue warps (ACT_B200_EW8=0 / 1) on the transformer's GEMM shapes,
graph-timed (steady state, L2-warm like inside the step): python scripts/ab_ew8.py"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
from scripts.kbench_graph import graph_time, bf

out = {"ew8": os.environ.get("ACT_B200_EW8", "1")}
for tag, M in (("enc", 3456), ("dec", 8192)):
    x384, x1536, x1152 = bf(M, 384), bf(M, 1536), bf(M, 1152)
    wqkv, wproj, w1, w2 = bf(1152, 384, scale=.05), bf(384, 384, scale=.05), bf(1536, 384, scale=.05), bf(384, 1536, scale=.05)
    b384, b1536 = torch.randn(384, device="cuda"), torch.randn(1536, device="cuda")
    xs = torch.randn(M, 384, device="cuda")
    u = torch.empty(M, 1536, dtype=torch.bfloat16, device="cuda")
    a = torch.empty(M, 1536, dtype=torch.bfloat16, device="cuda")
    o_qkv = torch.empty(M, 1152, dtype=torch.bfloat16, device="cuda")
    o_f32 = torch.empty(M, 384, device="cuda")
    o384 = torch.empty(M, 384, dtype=torch.bfloat16, device="cuda")
    g1536 = torch.zeros(1536, 384, device="cuda"); g384 = torch.zeros(384, 1536, device="cuda")
    cases = {
        "qkv_fwd": lambda: ops.gemm(x384, wqkv, out=o_qkv),
        "proj_fwd": lambda: ops.gemm(x384, wproj, bias=b384, resid=xs, out=o_f32),
        "fc1_fwd_gelu": lambda: ops.gemm(x384, w1, bias=b1536, act=1, preact_out=u, out=a),
        "fc2_fwd": lambda: ops.gemm(x1536, w2, bias=b384, resid=xs, out=o_f32),
        "fc2_dgrad_gelugrad": lambda: ops.gemm(x384, w2, b_mn=True, mul_in=u, mul_mode=1, out=a),
        "fc1_dgrad": lambda: ops.gemm(x1536, w1, b_mn=True, out=o384),
        "qkv_dgrad": lambda: ops.gemm(x1152, wqkv, b_mn=True, out=o384),
        "proj_dgrad": lambda: ops.gemm(x384, wproj, b_mn=True, out=o384),
        "fc1_wgrad": lambda: ops.gemm(x1536, x384, a_mn=True, b_mn=True, out=g1536, splits=5),
        "fc2_wgrad": lambda: ops.gemm(x384, x1536, a_mn=True, b_mn=True, out=g384, splits=5),
    }
    for name, fn in cases.items():
        try:
            out[f"{tag}_{name}"] = round(graph_time(fn), 2)
        except Exception as e:
            out[f"{tag}_{name}"] = str(e)[:60]
print(json.dumps(out))
