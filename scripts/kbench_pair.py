"""CTA-pair (cta_group::2) vs single-CTA GEMM kernels on the step's large K-major shapes (graph-replayed timing)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
dev = "cuda"
torch.manual_seed(0)
bf = lambda *s: (torch.randn(*s, device=dev) * 0.3).bfloat16()


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


shapes = [("vit_qkv_tok", 8192, 2304, 768, "bias"), ("vit_kv_prm", 8192, 1536, 768, "bias"), ("vit_proj", 8192, 768, 768, "resid"),
          ("vit_fc1", 8192, 3072, 768, "gelu"), ("vit_fc2", 8192, 768, 3072, "resid"), ("logits", 8192, 8192, 2304, "plain"),
          ("pn_conv3b", 262144, 512, 256, "plain"), ("pn_conv4", 106496, 384, 512, "bias"), ("dec_fc1", 8192, 1536, 384, "gelu"),
          ("square4k", 4096, 4096, 4096, "plain"), ("square8k", 8192, 8192, 8192, "plain")]
for name, M, N, K, mode in shapes:
    a, w = bf(M, K), bf(N, K)
    bias = torch.randn(N, device=dev)
    x = torch.randn(M, N, device=dev) if mode == "resid" else None
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if mode == "resid" else torch.bfloat16)
    kw = {"plain": {}, "bias": {"bias": bias}, "gelu": {"bias": bias, "act": ops.ACT_GELU}, "resid": {"bias": bias, "resid": x}}[mode]
    fl = 2.0 * M * N * K
    t1 = timeit(lambda: ops.gemm(a, w, out=out, **kw))
    t2 = timeit(lambda: ops.gemm(a, w, out=out, persistent=2, **kw))
    print(f"{name:12s} {M}x{N}x{K} {mode:5s}: auto {t1:7.1f} us ({fl / t1 / 1e6:6.0f} TF)   pair {t2:7.1f} us ({fl / t2 / 1e6:6.0f} TF)", flush=True)
