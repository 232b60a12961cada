"""Times the Stage-I dVAE training step (BASELINE config 3: N=1024, G=64, k=32, dims 256, 8192 tokens, B=64 =
cfgs/autoencoder/pointbert_dvae.yaml total_bs) on one GPU: eager and CUDA-graph, CUDA events, L2 flushed between steps.
Usage: python scripts/bench_dvae.py [B] [steps] ; prints one JSON object."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import dvae, engine, layers, ops
from act_b200.models import Cfg
from oracle import ref_model   # synthetic clouds + deterministic weights only (bench input generation)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg = Cfg(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256, tokens_dims=256, decoder_dims=256)
torch.cuda.set_device(0)
out = {"B": B, "steps": steps}
for use_graph in (False, True):
    model = ref_model.fill_params(dvae.DiscreteVAE(cfg), seed=8).cuda().train()
    fp = layers.FlatParams(model, lr=5e-4, weight_decay=5e-4)
    step = engine.AutoencoderStep(model, fp, B, 1024, use_graph=use_graph).capture()
    pts = ref_model.synthetic_clouds(B, 1024, seed=1).cuda()
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
print(json.dumps(out))
