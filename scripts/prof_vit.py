"""ncu target: one teacher ViT-B block as the step runs it (B=128 clouds, 64 token rows + 64 prompt rows as keys/values,
d=768): fused entry LayerNorm, q/k/v of the tokens, k/v of the prompts, prefix attention, proj(+resid), norm2, fc1(GELU),
fc2(+resid).
  ncu --set full --clock-control none --profile-from-start off -o /tmp/vit python scripts/prof_vit.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
B, G, P, D, H = 128, 64, 64, 768, 12
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(B * G, D, device=dev)
pos_tok = torch.randn(B * G, D, device=dev)
tok, ppos = torch.randn(P, D, device=dev), torch.randn(P, D, device=dev)
wqkv = (torch.randn(3 * D, D, device=dev) * .03).bfloat16()
wproj = (torch.randn(D, D, device=dev) * .03).bfloat16()
wfc1 = (torch.randn(4 * D, D, device=dev) * .03).bfloat16()
wfc2 = (torch.randn(D, 4 * D, device=dev) * .03).bfloat16()
b3 = torch.randn(3 * D, device=dev); b1 = torch.randn(D, device=dev); b4 = torch.randn(4 * D, device=dev)
g = torch.ones(D, device=dev); be = torch.zeros(D, device=dev)
seed = torch.tensor([7], dtype=torch.int64, device=dev)


def block():
    xs, h_tok, h_prm = ops.vit_ln1_fwd(x, pos_tok, tok, ppos, g, be, 1e-6, B, G, P, seed=seed, draw_id=1, p_drop=0.1)
    qkv_t = ops.gemm(h_tok, wqkv, bias=b3)
    kv_p = ops.gemm(h_prm, wqkv[D:], bias=b3[D:])
    o = ops.attention_prefix_fwd(qkv_t, kv_p, B, G, P, H, 0.125)
    xm = ops.gemm(o, wproj, bias=b1, resid=xs, out_dtype=torch.float32)
    h2, _, _, _ = ops.layernorm_fwd(xm, g, be, 1e-6, save_stats=False)
    a = ops.gemm(h2, wfc1, bias=b4, act=ops.ACT_GELU)
    return ops.gemm(a, wfc2, bias=b1, resid=xm, out_dtype=torch.float32)


for _ in range(3):
    block()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
block()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
