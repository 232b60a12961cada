"""Bring-up check of the CTA-pair (cta_group::2) GEMM kernel: python scripts/test_pair.py (run under `timeout`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
torch.manual_seed(0)
dev = "cuda"


def run(M, N, K, mode):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    lin = a.float() @ w.float().t()
    if mode == "plain":
        got = ops.gemm(a, w, out_dtype=torch.float32, persistent=2)
        want = lin
    elif mode == "gelu":
        got = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, persistent=2).float()
        want = torch.nn.functional.gelu(lin + bias)
    else:
        x = torch.randn(M, N, device=dev)
        got = ops.gemm(a, w, bias=bias, resid=x, out_dtype=torch.float32, persistent=2)
        want = lin + bias + x
    torch.cuda.synchronize()
    err = (got - want).abs().max().item() / (want.abs().max().item() + 1e-6)
    print(f"pair {mode:6s} {M}x{N}x{K}: max rel err {err:.2e}", flush=True)
    return err


bad = 0
for shape in [(256, 256, 64), (256, 256, 256), (512, 512, 768), (8192, 768, 768), (8192, 3072, 768), (8100, 2304, 768), (300, 264, 200)]:
    for mode in ("plain", "gelu", "resid"):
        e = run(*shape, mode)
        bad += e > (1e-2 if mode == "gelu" else 1e-4)
print("FAILED" if bad else "ALL OK")
