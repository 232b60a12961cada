"""One profiled Stage-I dVAE step for ncu: warm up, then bracket ONE eager step with cudaProfilerStart/Stop.
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/launches_dvae.csv python scripts/profile_dvae.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import data, dvae, engine, layers  # noqa: E402
from act_b200.models import Cfg  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
cfg = Cfg(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256, tokens_dims=256, decoder_dims=256)
model = dvae.DiscreteVAE(cfg).cuda().train()
fp = layers.FlatParams(model, lr=5e-4, weight_decay=5e-4)
step = engine.AutoencoderStep(model, fp, B, 1024, use_graph=False).capture()
pts = data.synthetic_clouds(B, 1024, seed=1).cuda()
for _ in range(2):
    step.run(pts)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step.run(pts)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
