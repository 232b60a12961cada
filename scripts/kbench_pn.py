"""Ablations of the pointnet GEMM shapes on the persistent kernel (graph-timed)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
from scripts.kbench_graph import graph_time, bf

M, BG = 262144, 8192
f2, a3 = bf(M, 256), bf(M, 512)
w3 = bf(512, 512, scale=.05)
w3c = w3[:, 256:].contiguous()
gpart = torch.randn(BG, 512, device="cuda")
o512 = torch.empty(M, 512, dtype=torch.bfloat16, device="cuda")
o256 = torch.empty(M, 256, dtype=torch.bfloat16, device="cuda")
o512f = torch.empty(M, 512, dtype=torch.float32, device="cuda")
cases = {
    "conv3 (view B, resid) bn256": lambda: ops.gemm(f2, w3[:, 256:], resid=gpart, resid_row_div=32, out=o512),
    "conv3 (view B, no resid) bn256": lambda: ops.gemm(f2, w3[:, 256:], out=o512),
    "conv3 (contig B, no resid) bn256": lambda: ops.gemm(f2, w3c, out=o512),
    "conv3 (contig B, no resid) bn128": lambda: ops.gemm(f2, w3c, out=o512, block_n=128),
    "conv3 no_out-ish: gmax only bn128": lambda: ops.gemm(f2, w3c, gmax_f32=torch.empty(BG, 512, device="cuda"), no_out=True),
    "K=512 N=256 Kmajor (a3 x w[256,512]) bn256": lambda: ops.gemm(a3, w3[:256, :], out=o256),
    "K=512 N=512 Kmajor bn256": lambda: ops.gemm(a3, w3, out=o512),
    "K=256 N=256 bn256": lambda: ops.gemm(f2, w3c[:256], out=o256),
    "conv3 v1 kernel bn128": lambda: ops.gemm(f2, w3c, out=o512, persistent=0),
}
for k, fn in cases.items():
    print(k, round(graph_time(fn, n=5), 1), flush=True)
