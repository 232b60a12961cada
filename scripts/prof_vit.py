"""ncu target: the teacher ViT-B's per-block kernels on their real shapes (B=128 clouds x 128 tokens, d=768).
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/vit python scripts/prof_vit.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
B, T, D, H = 128, 128, 768, 12
M = B * T
dev = "cuda"
x = torch.randn(M, D, device=dev)
h = torch.randn(M, D, device=dev).bfloat16()
a4 = torch.randn(M, 4 * D, device=dev).bfloat16()
wqkv = (torch.randn(3 * D, D, device=dev) * .03).bfloat16()
wproj = (torch.randn(D, D, device=dev) * .03).bfloat16()
wfc1 = (torch.randn(4 * D, D, device=dev) * .03).bfloat16()
wfc2 = (torch.randn(D, 4 * D, device=dev) * .03).bfloat16()
b3 = torch.randn(3 * D, device=dev); b1 = torch.randn(D, device=dev); b4 = torch.randn(4 * D, device=dev)
g = torch.ones(D, device=dev); be = torch.zeros(D, device=dev)


def block():
    h1, xs, _, _ = ops.layernorm_fwd(x, g, be, 1e-6, pos=x, save_stats=False)
    qkv = ops.gemm(h, wqkv, bias=b3)
    o, _ = ops.attention_fwd(qkv, B, T, H, 0.125)
    xm = ops.gemm(o, wproj, bias=b1, resid=x, out_dtype=torch.float32)
    a = ops.gemm(h, wfc1, bias=b4, act=ops.ACT_GELU)
    y = ops.gemm(a4, wfc2, bias=b1, resid=x, out_dtype=torch.float32)
    return y


for _ in range(3):
    block()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
block()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
