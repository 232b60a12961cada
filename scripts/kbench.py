"""Micro-benchmarks of individual act_b200 kernels (CUDA events, L2 flushed between iterations).
Usage (GPU box): python scripts/kbench.py [tokenizer|gemm|all] -> prints a table, writes gpurun_out/kbench.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops  # noqa: E402
from act_b200.data import synthetic_clouds  # noqa: E402

PEAKS = {"hbm_gbs": 6541.5, "bf16_tflops": 1639.0}
try:
    PEAKS.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def tokenizer(res):
    for (B, N, G, K) in [(128, 1024, 64, 32), (16, 8192, 512, 32)]:
        xyz = synthetic_clouds(B, N).cuda()
        med, best = timeit(lambda: ops.furthest_point_sample(xyz, G, return_center=True))
        by = B * (12 * N + 16 * G)
        res.append(dict(kernel="fps", shape=[B, N, G], us=med * 1e6, best_us=best * 1e6, gbs=by / med / 1e9,
                        us_per_round=med * 1e6 / (G - 1)))
        _, center = ops.furthest_point_sample(xyz, G, return_center=True)
        med, best = timeit(lambda: ops.knn(xyz, center, K, want_dist=False, want_neighborhood=True))
        by = B * (12 * N + 12 * G + 20 * G * K)
        res.append(dict(kernel="knn_group", shape=[B, N, G, K], us=med * 1e6, best_us=best * 1e6, gbs=by / med / 1e9))
    for (B, n, m) in [(4096, 8, 32), (4096, 32, 32), (1, 2048, 1024)]:
        a, b = torch.randn(B, n, 3, device="cuda"), torch.randn(B, m, 3, device="cuda")
        med, best = timeit(lambda: ops.chamfer_forward(a, b))
        by = B * (n + m) * 20
        res.append(dict(kernel="chamfer_fwd", shape=[B, n, m], us=med * 1e6, best_us=best * 1e6, gbs=by / med / 1e9))


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    res = []
    if what in ("tokenizer", "all"):
        tokenizer(res)
    if what in ("gemm", "all"):
        try:
            from scripts import kbench_gemm
            kbench_gemm.run(res, timeit, PEAKS)
        except ImportError:
            pass
    for r in res:
        print(json.dumps(r))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kbench.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
