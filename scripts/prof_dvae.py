"""Top kernels of the eager Stage-I dVAE step by device time (torch.profiler / CUPTI).  python scripts/prof_dvae.py [B]"""
import os, sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import data, dvae, engine, layers
from act_b200.models import Cfg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = Cfg(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256, tokens_dims=256, decoder_dims=256)
torch.manual_seed(0)
model = dvae.DiscreteVAE(cfg).cuda().train()
fp = layers.FlatParams(model, lr=5e-4, weight_decay=5e-4)
step = engine.AutoencoderStep(model, fp, B, 1024, use_graph=False).capture()
pts = data.synthetic_clouds(B, 1024, seed=1).cuda()
for _ in range(2):
    step.run(pts)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step.run(pts)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
