"""Tile / split-K sweep of the weight-gradient GEMMs dW[N,K] = dY^T X (both operands MN-major, fp32 atomics), graph-timed:
python scripts/sweep_wgrad_tiles.py"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
from scripts.kbench_graph import graph_time, bf

for T in (3456, 8192):
    for (N, K) in ((1536, 384), (384, 1536), (1152, 384), (384, 384)):
        dy, x = bf(T, N), bf(T, K)
        g = torch.zeros(N, K, device="cuda")
        row = {"default_splits": ops.wgrad_splits(N, K, T)}
        row["default"] = round(graph_time(lambda: ops.wgrad(dy, x, g)), 2)
        for bn in (64, 128, 192):
            if K % bn and bn != 128:
                continue
            for sp in (2, 3, 4, 5, 6, 8, 12, 17):
                try:
                    row[f"bn{bn}_s{sp}"] = round(graph_time(lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=g, splits=sp, block_n=bn, persistent=0)), 2)
                except Exception as e:
                    row[f"bn{bn}_s{sp}"] = "x"
        best = min((v, k) for k, v in row.items() if isinstance(v, float) and k != "default")
        print(f"T={T} dW[{N},{K}] default {row['default']} (splits {row['default_splits']}) best {best}", json.dumps(row), flush=True)
