"""Epilogue experiments: same GEMM shape under different epilogue modes / tile widths (graph-replayed timing)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
dev = "cuda"
torch.manual_seed(0)
bf = lambda *s: (torch.randn(*s, device=dev) * 0.5).bfloat16()


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e3


M = 262144
a, w = bf(M, 256), bf(512, 256)
res = torch.randn(M // 32, 512, device=dev)
ob = torch.empty(M, 512, dtype=torch.bfloat16, device=dev)
of = torch.empty(M // 4, 512, dtype=torch.float32, device=dev)
b = torch.randn(512, device=dev)
for name, fn in {
    "conv3a plain bf16 p256": lambda: ops.gemm(a, w, out=ob),
    "conv3a plain bf16 p128": lambda: ops.gemm(a, w, out=ob, block_n=128, persistent=1),
    "conv3a plain bf16 onetile128": lambda: ops.gemm(a, w, out=ob, block_n=128, persistent=0),
    "conv3a +bias bf16 p256": lambda: ops.gemm(a, w, out=ob, bias=b),
    "conv3a +resid bcast bf16 p256": lambda: ops.gemm(a, w, out=ob, resid=res, resid_row_div=32),
    "conv3a +resid bcast bf16 p128": lambda: ops.gemm(a, w, out=ob, resid=res, resid_row_div=32, block_n=128, persistent=1),
    "conv3a(M/4) plain f32 p256": lambda: ops.gemm(a[:M // 4], w, out=of),
}.items():
    print(f"{name:40s} {timeit(fn):8.1f} us", flush=True)
M2 = 16384
a2, w2 = bf(M2, 768), bf(768, 768)
r2, b2 = torch.randn(M2, 768, device=dev), torch.randn(768, device=dev)
o2 = torch.empty(M2, 768, device=dev)
o2b = torch.empty(M2, 768, device=dev, dtype=torch.bfloat16)
for name, fn in {
    "vit_proj plain f32 p256": lambda: ops.gemm(a2, w2, out=o2),
    "vit_proj plain bf16 p256": lambda: ops.gemm(a2, w2, out=o2b),
    "vit_proj +bias+resid f32 p256": lambda: ops.gemm(a2, w2, out=o2, bias=b2, resid=r2),
    "vit_proj +bias+resid f32 p128": lambda: ops.gemm(a2, w2, out=o2, bias=b2, resid=r2, block_n=128, persistent=1),
    "vit_proj +bias+resid f32 onetile128": lambda: ops.gemm(a2, w2, out=o2, bias=b2, resid=r2, block_n=128, persistent=0),
    "vit_proj +bias+resid f32 onetile192": lambda: ops.gemm(a2, w2, out=o2, bias=b2, resid=r2, block_n=192, persistent=0),
}.items():
    print(f"{name:40s} {timeit(fn):8.1f} us", flush=True)
