"""A/B of the FPS kernels in the dense regime (CUDA events, L2 flushed): run twice, ACT_B200_FPS_CLUSTER=0 and default."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import data, ops
from scripts.kbench import timeit
res = []
for (B, N, G) in [(16, 8192, 512), (16, 4096, 256), (2, 8192, 512)]:
    xyz = data.synthetic_clouds(B, N).cuda()
    med, best = timeit(lambda: ops.furthest_point_sample(xyz, G, return_center=True))
    res.append(dict(kernel="fps", cluster=os.environ.get("ACT_B200_FPS_CLUSTER", "1"), shape=[B, N, G], us=round(med * 1e6, 1),
                    us_per_round=round(med * 1e6 / (G - 1), 3), gbs=round(B * (12 * N + 16 * G) / med / 1e9, 2)))
for (B, N, G, K) in [(16, 8192, 512, 32), (16, 4096, 256, 32), (128, 1024, 64, 32)]:
    xyz = data.synthetic_clouds(B, N).cuda()
    _, center = ops.furthest_point_sample(xyz, G, return_center=True)
    med, best = timeit(lambda: ops.knn(xyz, center, K, want_dist=False, want_neighborhood=True))
    res.append(dict(kernel="knn_group", shape=[B, N, G, K], us=round(med * 1e6, 1),
                    gbs=round(B * (12 * N + 12 * G + 20 * G * K) / med / 1e9, 2)))
for r in res:
    print(json.dumps(r))
