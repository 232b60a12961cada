"""Per-call time of attention forward / backward (CUDA graph of 20 calls, CUDA events): run once with ACT_B200_ATTN_TC=1 and
once with =0 to compare the tcgen05 kernels with the warp-MMA / FMA ones."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from act_b200 import ops

def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return best * 1e3

out = {"tc": os.environ.get("ACT_B200_ATTN_TC", "1")}
for B, T, H in [(128, 27, 6), (128, 64, 6), (128, 65, 6), (16, 206, 6), (16, 512, 6), (128, 128, 12)]:
    qkv = (torch.randn(B * T, 3 * H * 64, device="cuda") * 0.8).bfloat16()
    do = (torch.randn(B * T, H * 64, device="cuda") * 0.5).bfloat16()
    o, lse = ops.attention_fwd(qkv, B, T, H, 0.125)
    f = timeit(lambda: ops.attention_fwd(qkv, B, T, H, 0.125))
    b = timeit(lambda: ops.attention_bwd(qkv, o, do, lse, B, T, H, 0.125))
    out[f"B{B}_T{T}_H{H}"] = {"fwd_us": round(f, 2), "bwd_us": round(b, 2)}
print(json.dumps(out))
