"""Warm / L2-flushed timings of the small bandwidth kernels of the Stage-II step (CUDA events): python scripts/kbench_misc.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
from scripts.kbench import timeit

res = []
def rec(name, fn, nbytes):
    med, best = timeit(fn)
    res.append(dict(kernel=name, us=round(med * 1e6, 1), best_us=round(best * 1e6, 1), gbs=round(nbytes / med / 1e9, 1)))

for R in (8192, 3456):
    da = torch.randn(R, 128, device="cuda").bfloat16(); x = torch.randn(R, 3, device="cuda")
    W = torch.randn(128, 3, device="cuda"); b = torch.randn(128, device="cuda")
    dW = torch.zeros(128, 3, device="cuda"); db = torch.zeros(128, device="cuda")
    rec(f"pos_mlp1_bwd R={R}", lambda: ops.pos_mlp1_bwd(da, x, W, b, dW, db), R * 128 * 2)
for (M, C) in ((3456, 1536), (8192, 1536), (3456, 384)):
    xx = torch.randn(M, C, device="cuda").bfloat16(); out = torch.zeros(C, device="cuda")
    rec(f"colsum {M}x{C}", lambda: ops.colsum(xx, out), M * C * 2)
B, G, P, H = 128, 64, 64, 12
qkv = torch.randn(B * G, 3 * H * 64, device="cuda").bfloat16(); kvp = torch.randn(B * P, 2 * H * 64, device="cuda").bfloat16()
rec("attn_prefix_fwd", lambda: ops.attention_prefix_fwd(qkv, kvp, B, G, P, H, 0.125), (qkv.numel() + kvp.numel() + B * G * H * 64) * 2)
C = 768
x = torch.randn(B * G, C, device="cuda"); pos = torch.randn(B * G, C, device="cuda")
tok = torch.randn(P, C, device="cuda"); ppos = torch.randn(P, C, device="cuda")
g = torch.ones(C, device="cuda"); be = torch.zeros(C, device="cuda")
seed = torch.tensor([5], dtype=torch.int64, device="cuda")
rec("vit_ln1", lambda: ops.vit_ln1_fwd(x, pos, tok, ppos, g, be, 1e-6, B, G, P, seed=seed, draw_id=1, p_drop=0.1),
    B * G * C * (4 + 4 + 4 + 2) + B * P * C * 2)
for (M, C) in ((3456, 384), (8192, 384), (8192, 768)):
    xx = torch.randn(M, C, device="cuda"); gg = torch.ones(C, device="cuda"); bb = torch.zeros(C, device="cuda")
    rec(f"layernorm_fwd {M}x{C}", lambda: ops.layernorm_fwd(xx, gg, bb), M * C * 6)
for r in res:
    print(json.dumps(r))
