"""One launch each of the conv4-like fused-max GEMM (262144x384x512) and conv3-like (262144x512x256) for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
M = 262144
a = (torch.randn(M, 512, device="cuda") * 0.5).bfloat16(); w = (torch.randn(384, 512, device="cuda") * 0.05).bfloat16()
bias = torch.randn(384, device="cuda"); gf = torch.empty(M // 32, 384, device="cuda")
a2 = (torch.randn(M, 128, device="cuda") * 0.5).bfloat16(); w2 = (torch.randn(256, 128, device="cuda") * 0.05).bfloat16()
b2 = torch.randn(256, device="cuda"); gb = torch.empty(M // 32, 256, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ops.gemm(a, w, bias=bias, gmax_f32=gf, no_out=True)
    ops.gemm(a2, w2, bias=b2, gmax_bf16=gb)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ops.gemm(a, w, bias=bias, gmax_f32=gf, no_out=True)
ops.gemm(a2, w2, bias=b2, gmax_bf16=gb)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
