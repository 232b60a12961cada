"""A/B of the fused max-over-32-rows GEMMs (mini-PointNet conv2 / conv4): run with ACT_B200_PAIR_GMAX=0 / 1 / 2."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
from scripts.kbench import timeit
for (M, N, K, f32) in [(262144, 384, 512, True), (106496, 384, 512, True), (131072, 256, 512, True), (262144, 256, 128, False)]:
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    G = M // 32
    if f32:
        gf = torch.empty(G, N, device="cuda")
        fn = lambda: ops.gemm(a, w, bias=bias, gmax_f32=gf, no_out=True)
    else:
        gb = torch.empty(G, N, dtype=torch.bfloat16, device="cuda")
        fn = lambda: ops.gemm(a, w, bias=bias, gmax_bf16=gb)
    med, best = timeit(fn)
    print(json.dumps(dict(mode=os.environ.get("ACT_B200_PAIR_GMAX", "default"), shape=[M, N, K], us=round(med * 1e6, 1),
                          tflops=round(2.0 * M * N * K / med / 1e12, 1))))
