"""A/B of the frozen teacher's DGCNN edge kernel (dgcnn_edge_gn) between two builds of libact_b200.so:
    python scripts/ab_edge.py                      # times the in-tree library
    ACT_B200_LIB=path/to/other.so python scripts/ab_edge.py
L2-flushed CUDA-event medians at the Stage-II shapes (B=128, G=64, the four DGCNN widths), plus a checksum of the output
so that the two builds can be compared bit for bit."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops, _lib
from scripts.kbench import timeit

torch.manual_seed(0)
B, G = 128, 64
coor = torch.randn(B, G, 3, device="cuda")
_, idx, _ = ops.knn(coor, coor, 4, want_dist=False)
out = {"lib": _lib.LIB_PATH}
for Cp in (256, 512, 1024):
    pq = torch.randn(B * G, 2 * Cp, device="cuda")
    gam, bet = torch.rand(Cp, device="cuda") + 0.5, torch.randn(Cp, device="cuda") * 0.1
    feats = torch.zeros(B * G, 2304, dtype=torch.bfloat16, device="cuda")
    f = feats[:, 256:256 + Cp]
    med, best = timeit(lambda: ops.dgcnn_edge_gn(pq, idx, gam, bet, B, G, Cp, 1e-5, 0.2, f), iters=40)
    out[f"Cp{Cp}"] = dict(us=round(med * 1e6, 2), best_us=round(best * 1e6, 2),
                          checksum=float(f.float().double().sum().item()), absum=float(f.float().abs().double().sum().item()))
print(json.dumps(out))
