"""Times the Stage-I dVAE training step (BASELINE config 3: N=1024, G=64, k=32, dims 256, 8192 tokens, B=64 =
cfgs/autoencoder/pointbert_dvae.yaml total_bs) on one GPU: eager and CUDA-graph, CUDA events, L2 flushed between steps.
Usage: python scripts/bench_dvae.py [B] [steps] ; prints one JSON object."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import data, dvae, engine, layers, ops
from act_b200.models import Cfg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = Cfg(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256, tokens_dims=256, decoder_dims=256)
torch.cuda.set_device(0)
out = {"B": B, "steps": steps}
for use_graph in (False, True):
    torch.manual_seed(0)
    model = dvae.DiscreteVAE(cfg).cuda().train()
    fp = layers.FlatParams(model, lr=5e-4, weight_decay=5e-4)
    step = engine.AutoencoderStep(model, fp, B, 1024, use_graph=use_graph).capture()
    pts = data.synthetic_clouds(B, 1024, seed=1).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        step.run(pts)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = time.perf_counter()
    first = None
    for a, b in ev:
        flush.zero_()
        a.record()
        l = step.run(pts)
        b.record()
        if first is None:
            first = l.clone()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    out["graph" if use_graph else "eager"] = {"ms_per_step": round(ms, 3), "clouds_per_s": round(B / ms * 1e3, 1),
                                              "wall_ms_per_step": round(wall / steps * 1e3, 3),
                                              "act_b200_launches": step.launches_per_step,
                                              "loss_first": [round(x, 5) for x in first.tolist()],
                                              "loss_last": [round(x, 5) for x in l.tolist()]}
    del step, fp, model
    torch.cuda.empty_cache()


def cpu_baseline(batch=8):
    """The reference's Stage-I step restated for the host cores (oracle/ref_dvae.py + the C Group / Chamfer), all threads."""
    import os as _os
    from oracle import ref_dvae
    threads = _os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    m = ref_dvae.DiscreteVAE().train()
    opt = torch.optim.AdamW(m.parameters(), lr=5e-4, weight_decay=5e-4)
    p = data.synthetic_clouds(batch, 1024, seed=1)
    ts = []
    for i in range(2):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        ret = m(p, temperature=1.0, hard=False)
        l1, l2 = m.get_loss(ret, p)
        (l1 + 0.0 * l2).backward()
        opt.step()
        ts.append(time.perf_counter() - t0)
    return {"value": round(batch / ts[-1], 2), "unit": "clouds/s", "cores": threads, "kind": "port",
            "sample": f"1 timed step after 1 warm-up of oracle/ref_dvae.py at batch {batch}"}


if os.environ.get("DVAE_CPU_BASELINE", "1") != "0":
    out["cpu_baseline"] = cpu_baseline()
out["config"] = {"workload": "dVAE Stage-I step (BASELINE config 3): N=1024, G=64 x k=32, dims 256, 8192 tokens, "
                             "fwd + ChamferL1 x2 + KL + bwd + AdamW", "batch": B,
                 "l2": "256 MB L2-flush write between timed steps, outside the event pairs"}
print(json.dumps(out))
