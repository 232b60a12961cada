import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
M = 262144
f2 = (torch.randn(M, 256, device="cuda")).bfloat16()
w = (torch.randn(512, 256, device="cuda") * .05).bfloat16()
o = torch.empty(M, 512, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ops.gemm(f2, w, out=o)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ops.gemm(f2, w, out=o)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
