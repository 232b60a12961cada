"""Per-launch steady-state kernel times: N back-to-back launches of one op captured in a CUDA graph, replayed and
timed with CUDA events (no host launch overhead in the number; operands stay L2-warm like inside the real step).
Usage (GPU box): python scripts/kbench_graph.py  -> table + gpurun_out/kbench_graph.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import ops  # noqa: E402

PEAK = 1639.0
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass


def graph_time(fn, n=20, reps=5):
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / n)
    return best


def bf(*s, scale=1.0):
    return (torch.randn(*s, device="cuda") * scale).bfloat16()


def main():
    res = []

    def rec(name, us, flops=None, bytes_=None):
        r = {"kernel": name, "us": round(us, 2)}
        if flops:
            r["tflops"] = round(flops / us / 1e6, 1)
            r["frac_peak"] = round(flops / us / 1e6 / PEAK, 3)
        if bytes_:
            r["gbs"] = round(bytes_ / us / 1e3, 1)
        res.append(r)
        print(json.dumps(r), flush=True)

    # transformer GEMMs (encoder M = 128*27, decoder M = 128*64)
    for tag, M in (("enc", 3456), ("dec", 8192)):
        x384, x1536 = bf(M, 384), bf(M, 1536)
        wqkv, wproj, w1, w2 = bf(1152, 384, scale=.05), bf(384, 384, scale=.05), bf(1536, 384, scale=.05), bf(384, 1536, scale=.05)
        b384, b1536 = torch.randn(384, device="cuda"), torch.randn(1536, device="cuda")
        xs = torch.randn(M, 384, device="cuda")
        o_qkv = torch.empty(M, 1152, dtype=torch.bfloat16, device="cuda")
        o_f32 = torch.empty(M, 384, device="cuda")
        u = torch.empty(M, 1536, dtype=torch.bfloat16, device="cuda")
        a = torch.empty(M, 1536, dtype=torch.bfloat16, device="cuda")
        o384 = torch.empty(M, 384, dtype=torch.bfloat16, device="cuda")
        gW = torch.zeros(1536, 384, device="cuda")
        gW2 = torch.zeros(384, 1536, device="cuda")
        rec(f"{tag}_qkv_fwd", graph_time(lambda: ops.gemm(x384, wqkv, out=o_qkv)), 2.0 * M * 1152 * 384)
        rec(f"{tag}_proj_fwd+bias+resid", graph_time(lambda: ops.gemm(x384, wproj, bias=b384, resid=xs, out=o_f32)), 2.0 * M * 384 * 384)
        rec(f"{tag}_fc1_fwd+gelu", graph_time(lambda: ops.gemm(x384, w1, bias=b1536, act=1, preact_out=u, out=a)), 2.0 * M * 1536 * 384)
        rec(f"{tag}_fc2_fwd+bias+resid", graph_time(lambda: ops.gemm(x1536, w2, bias=b384, resid=xs, out=o_f32)), 2.0 * M * 1536 * 384)
        rec(f"{tag}_fc2_dgrad+gelugrad", graph_time(lambda: ops.gemm(x384, w2, b_mn=True, mul_in=u, mul_mode=1, out=a)), 2.0 * M * 1536 * 384)
        rec(f"{tag}_fc1_dgrad", graph_time(lambda: ops.gemm(x1536, w1, b_mn=True, out=o384)), 2.0 * M * 1536 * 384)
        rec(f"{tag}_fc1_wgrad", graph_time(lambda: ops.wgrad(x1536, x384, gW)), 2.0 * M * 1536 * 384)
        rec(f"{tag}_fc2_wgrad", graph_time(lambda: ops.wgrad(x384, x1536, gW2)), 2.0 * M * 1536 * 384)
        T = 27 if tag == "enc" else 64
        B = M // T
        qkv = bf(M, 1152, scale=.5)
        o, lse = ops.attention_fwd(qkv, B, T, 6, 0.125)
        do = bf(M, 384)
        fl = 4.0 * B * 6 * T * T * 64
        rec(f"{tag}_attn_fwd_T{T}", graph_time(lambda: ops.attention_fwd(qkv, B, T, 6, 0.125)), fl)
        rec(f"{tag}_attn_bwd_T{T}", graph_time(lambda: ops.attention_bwd(qkv, o, do, lse, B, T, 6, 0.125)), 2.5 * fl)
        g_, b_ = torch.ones(384, device="cuda"), torch.zeros(384, device="cuda")
        y, _, mean, rstd = ops.layernorm_fwd(xs, g_, b_, pos=xs)
        rec(f"{tag}_ln_fwd", graph_time(lambda: ops.layernorm_fwd(xs, g_, b_, pos=xs)), None, M * 384 * (4 + 4 + 4 + 2))
        dg, db, dacc = torch.zeros(384, device="cuda"), torch.zeros(384, device="cuda"), torch.zeros(M, 384, device="cuda")
        rec(f"{tag}_ln_bwd", graph_time(lambda: ops.layernorm_bwd(o384, xs, mean, rstd, g_, dg, db, dres=xs, dacc=dacc, want_bf16=True, dbias=db)),
            None, M * 384 * (2 + 4 + 4 + 4 + 8 + 2))
    # mini-PointNet GEMMs, in-model form
    M, BG = 262144, 8192
    a1, f2, a3 = bf(M, 128), bf(M, 256), bf(M, 512)
    w2, w3, w4 = bf(256, 128, scale=.05), bf(512, 512, scale=.05), bf(384, 512, scale=.05)
    b256, b384 = torch.randn(256, device="cuda"), torch.randn(384, device="cuda")
    gpart = torch.randn(BG, 512, device="cuda")
    o256 = torch.empty(M, 256, dtype=torch.bfloat16, device="cuda")
    o512 = torch.empty(M, 512, dtype=torch.bfloat16, device="cuda")
    gm = torch.empty(BG, 256, dtype=torch.bfloat16, device="cuda")
    ga = torch.empty(BG, 256, dtype=torch.uint8, device="cuda")
    tok = torch.empty(BG, 384, device="cuda")
    ga4 = torch.empty(BG, 384, dtype=torch.uint8, device="cuda")
    dF4 = bf(M, 384)
    rec("pn_conv2_fwd+gmax", graph_time(lambda: ops.gemm(a1, w2, bias=b256, gmax_bf16=gm, garg=ga, out=o256), n=5), 2.0 * M * 256 * 128, M * (128 + 256) * 2)
    rec("pn_conv3_fwd(K=256)+gpart", graph_time(lambda: ops.gemm(f2, w3[:, 256:], resid=gpart, resid_row_div=32, out=o512), n=5), 2.0 * M * 512 * 256, M * (256 + 512) * 2)
    rec("pn_conv4_fwd+gmax(no out)", graph_time(lambda: ops.gemm(a3, w4, bias=b384, gmax_f32=tok, garg=ga4, no_out=True), n=5), 2.0 * M * 384 * 512, M * 512 * 2)
    rec("pn_conv4_dgrad+relumask", graph_time(lambda: ops.gemm(dF4, w4, b_mn=True, mul_in=a3, mul_mode=2, out=o512), n=5), 2.0 * M * 384 * 512, M * (384 + 512 + 512) * 2)
    rec("pn_conv3_dgrad", graph_time(lambda: ops.gemm(a3, w3[:, 256:], b_mn=True, out=o256), n=5), 2.0 * M * 512 * 256, M * (512 + 256) * 2)
    gw4 = torch.zeros(384, 512, device="cuda")
    rec("pn_conv4_wgrad", graph_time(lambda: ops.wgrad(dF4, a3, gw4), n=5), 2.0 * M * 384 * 512, M * (384 + 512) * 2)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/kbench_graph.json", "w"), indent=1)


if __name__ == "__main__":
    main()
