"""ncu target: the teacher ViT's prefix attention (64 token queries x (64 prompt + 64 token) keys, 12 heads, 128 clouds)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops
B, G, P, H = 128, 64, 64, 12
C = H * 64
qkv_t = (torch.randn(B * G, 3 * C, device="cuda") * 0.5).bfloat16()
kv_p = (torch.randn(B * P, 2 * C, device="cuda") * 0.5).bfloat16()
for _ in range(3):
    ops.attention_prefix_fwd(qkv_t, kv_p, B, G, P, H, 0.125)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attention_prefix_fwd(qkv_t, kv_p, B, G, P, H, 0.125)
e1.record(); torch.cuda.synchronize()
print("prefix attention us/launch (back-to-back eager):", e0.elapsed_time(e1) / 20 * 1e3)
torch.cuda.cudart().cudaProfilerStart()
ops.attention_prefix_fwd(qkv_t, kv_p, B, G, P, H, 0.125)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
