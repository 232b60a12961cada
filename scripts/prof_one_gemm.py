"""ncu target: ONE GEMM call of the step on its real shape.
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/g python scripts/prof_one_gemm.py <case>"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
case = sys.argv[1] if len(sys.argv) > 1 else "conv3a"
dev = "cuda"
torch.manual_seed(0)
bf = lambda *s: (torch.randn(*s, device=dev) * 0.5).bfloat16()
if case == "conv3a":          # 262144x512x256 + per-group broadcast residual row, bf16 out
    M = 262144
    a, w = bf(M, 256), bf(512, 256)
    res = torch.randn(M // 32, 512, device=dev)
    o = torch.empty(M, 512, dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.gemm(a, w, out=o, resid=res, resid_row_div=32)
elif case == "vit_proj":
    M = 16384
    a, w = bf(M, 768), bf(768, 768)
    res, b = torch.randn(M, 768, device=dev), torch.randn(768, device=dev)
    o = torch.empty(M, 768, device=dev)
    fn = lambda: ops.gemm(a, w, out=o, bias=b, resid=res)
elif case == "enc_fc1":
    M = 3456
    a, w, b = bf(M, 384), bf(1536, 384), torch.randn(1536, device=dev)
    o = torch.empty(M, 1536, dtype=torch.bfloat16, device=dev)
    pre = torch.empty(M, 1536, dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.gemm(a, w, out=o, bias=b, act=ops.ACT_GELU, preact_out=pre)
elif case == "enc_proj":
    M = 3456
    a, w, b = bf(M, 384), bf(384, 384), torch.randn(384, device=dev)
    res = torch.randn(M, 384, device=dev)
    rs = torch.ones(128, device=dev)
    fn = lambda: ops.gemm(a, w, out=res, out_dtype=torch.float32, bias=b, resid=res, row_scale=rs, rows_per_scale=27)
else:
    raise SystemExit("unknown case")
for _ in range(3):
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
