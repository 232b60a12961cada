"""One profiled Stage-II step for ncu: warm up, then bracket ONE step (or N) with cudaProfilerStart/Stop.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python scripts/profile_step.py [steps] [batch] [native|synthetic]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import layers, models  # noqa: E402
from act_b200.data import synthetic_clouds  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
np.random.seed(0)
TEACHER = sys.argv[3] if len(sys.argv) > 3 else "native"
NP = int(sys.argv[4]) if len(sys.argv) > 4 else 1024          # points per cloud (8192 = dense regime)
G = int(sys.argv[5]) if len(sys.argv) > 5 else 64             # groups per cloud (512 = dense regime)
model = models.ACT_PointDistillation(models.default_config(0.6, 0.1, num_group=G),
                                     teacher="native" if TEACHER == "native" else "synthetic").cuda().train()
fp = layers.FlatParams(model, exclude=model.UNUSED_PARAMETERS)
pts = synthetic_clouds(B, NP).cuda()


def step():
    fp.zero_grad()
    loss = model(pts)
    loss.backward()
    fp.set_hyper()
    fp.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
