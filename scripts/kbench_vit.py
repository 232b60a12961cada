"""Steady-state per-launch times of the teacher ViT-B block kernels on their real shapes (B=128 x 128 tokens, d=768):
each call replayed 10x from a CUDA graph, CUDA events on the launching stream, L2 not flushed (operands >> L2 anyway).
  python scripts/kbench_vit.py [out.json]"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
B, T, D, H, P = 128, 128, 768, 12, 64
M = B * T
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(M, D, device=dev)
h = torch.randn(M, D, device=dev).bfloat16()
a4 = torch.randn(M, 4 * D, device=dev).bfloat16()
qkv = (torch.randn(M, 3 * D, device=dev) * .5).bfloat16()
o = torch.randn(M, D, device=dev).bfloat16()
wqkv = (torch.randn(3 * D, D, device=dev) * .03).bfloat16()
wproj = (torch.randn(D, D, device=dev) * .03).bfloat16()
wfc1 = (torch.randn(4 * D, D, device=dev) * .03).bfloat16()
wfc2 = (torch.randn(D, 4 * D, device=dev) * .03).bfloat16()
b3 = torch.randn(3 * D, device=dev); b1 = torch.randn(D, device=dev); b4 = torch.randn(4 * D, device=dev)
g = torch.ones(D, device=dev); be = torch.zeros(D, device=dev)
pos_tok = torch.randn(B * 64, D, device=dev)
tok = torch.randn(P, D, device=dev); ppos = torch.randn(P, D, device=dev)
seed = torch.tensor([7], dtype=torch.int64, device=dev)
out32 = torch.empty(M, D, device=dev)
outb = torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16)
outq = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
cases = {
    "vit_ln1_fused(prompt+pos+norm1)": (lambda: ops.vit_ln1_fwd(x[:B * 64], pos_tok, tok, ppos, g, be, 1e-6, B, 64, P, seed=seed, draw_id=1, p_drop=0.1), None),
    "vit_attn_prefix 64q x 128kv H=12": (lambda: ops.attention_prefix_fwd(qkv[:B * 64], qkv[B * 64:, :2 * D].contiguous(), B, 64, P, H, 0.125), 4.0 * B * H * 64 * T * 64),
    "vit_qkv_tok 8192x2304x768+bias": (lambda: ops.gemm(h[:B * 64], wqkv, bias=b3), 2.0 * B * 64 * 3 * D * D),
    "vit_kv_prm 8192x1536x768+bias": (lambda: ops.gemm(h[:B * 64], wqkv[D:], bias=b3[D:]), 2.0 * B * 64 * 2 * D * D),
    "vit_proj 8192x768x768+bias+resid": (lambda: ops.gemm(o[:B * 64], wproj, bias=b1, resid=x[:B * 64], out=out32[:B * 64]), 2.0 * B * 64 * D * D),
    "vit_fc1 8192x3072x768+bias+gelu": (lambda: ops.gemm(h[:B * 64], wfc1, bias=b4, act=ops.ACT_GELU, out=outb[:B * 64]), 2.0 * B * 64 * 4 * D * D),
    "vit_fc2 8192x768x3072+bias+resid": (lambda: ops.gemm(a4[:B * 64], wfc2, bias=b1, resid=x[:B * 64], out=out32[:B * 64]), 2.0 * B * 64 * 4 * D * D),
    "vit_qkv 16384x2304x768+bias": (lambda: ops.gemm(h, wqkv, bias=b3, out=outq), 2.0 * M * 3 * D * D),
    "vit_attn_fwd T=128 H=12": (lambda: ops.attention_fwd(qkv, B, T, H, 0.125), 4.0 * B * H * T * T * 64),
    "vit_proj 16384x768x768+bias+resid": (lambda: ops.gemm(o, wproj, bias=b1, resid=x, out=out32), 2.0 * M * D * D),
    "vit_ln2": (lambda: ops.layernorm_fwd(x, g, be, 1e-6, save_stats=False), None),
    "vit_fc1 16384x3072x768+bias+gelu": (lambda: ops.gemm(h, wfc1, bias=b4, act=ops.ACT_GELU, out=outb), 2.0 * M * 4 * D * D),
    "vit_fc2 16384x768x3072+bias+resid": (lambda: ops.gemm(a4, wfc2, bias=b1, resid=x, out=out32), 2.0 * M * 4 * D * D),
}
res = []
for name, (fn, fl) in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        for _ in range(10):
            fn()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gph.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    r = {"kernel": name, "us": round(best * 1e3, 2)}
    if fl:
        r["tflops"] = round(fl / (best * 1e-3) / 1e12, 1)
    res.append(r)
    print(r, flush=True)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
