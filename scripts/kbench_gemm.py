"""GEMM micro-benchmarks (called from scripts/kbench.py)."""
import torch

from act_b200 import ops


def run(res, timeit, peaks):
    shapes = [("enc_qkv", 3456, 1152, 384), ("enc_proj", 3456, 384, 384), ("enc_fc1", 3456, 1536, 384),
              ("enc_fc2", 3456, 384, 1536), ("pn_conv2", 262144, 256, 128), ("pn_conv3", 262144, 512, 512),
              ("pn_conv4", 262144, 384, 512), ("square4k", 4096, 4096, 4096)]
    for name, M, N, K in shapes:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        for bn in (64, 128):
            med, best = timeit(lambda: ops.gemm(a, b, out=out, block_n=bn), iters=10)
            fl = 2.0 * M * N * K
            res.append(dict(kernel=f"gemm_{name}_bn{bn}", shape=[M, N, K], us=med * 1e6, best_us=best * 1e6,
                            tflops=fl / med / 1e12, frac_of_measured_peak=fl / med / 1e12 / peaks["bf16_tflops"]))
        med, best = timeit(lambda: torch.matmul(a, b.t(), out=out), iters=10)
        res.append(dict(kernel=f"cublas_{name}", shape=[M, N, K], us=med * 1e6, tflops=2.0 * M * N * K / med / 1e12))
