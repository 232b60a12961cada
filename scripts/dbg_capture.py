import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from act_b200 import layers, models
from act_b200.engine import PretrainStep
from oracle import ref_model
torch.manual_seed(0); np.random.seed(0)
cfg = models.default_config(mask_ratio=0.6, drop_path_rate=0.0)
model = ref_model.fill_params(models.ACT_PointDistillation(cfg, teacher="synthetic"), seed=3).cuda().train()
fp = layers.FlatParams(model, lr=1e-3, exclude=model.UNUSED_PARAMETERS)
from act_b200 import ops
mode = os.environ.get("DBG", "")
pts = ref_model.synthetic_clouds(8, 1024, seed=1).cuda()
mask = modules_mask = None
from act_b200.modules import mask_center_rand
mask = mask_center_rand(8, 64, 0.6, "cuda")
torch.cuda.synchronize()
import inspect
def _wrap(cls):
    orig = cls.backward
    def bw(ctx, *a):
        print("  bwd enter", cls.__name__, "stream", hex(torch.cuda.current_stream().cuda_stream), flush=True)
        try:
            r = orig(ctx, *a)
        except Exception as e:
            print("  bwd RAISE", cls.__name__, type(e).__name__, str(e)[:120], flush=True)
            raise
        print("  bwd exit", cls.__name__, flush=True)
        return r
    cls.backward = staticmethod(bw)
if "w" in mode:
    for n, c in list(vars(layers).items()):
        if inspect.isclass(c) and issubclass(c, torch.autograd.Function) and c is not torch.autograd.Function:
            _wrap(c)
import torch.nn.functional as F
if "1" in mode:
    layers.pos_mlp = lambda seq, x: layers.linear(F.gelu(F.linear(x, seq[0].weight, seq[0].bias)), seq[2].weight, seq[2].bias)
if "2" in mode:
    def _asm(src, fill, B, n, T, ff, src_off=0):
        C = fill.numel()
        s3 = src.reshape(B, -1, C)[:, src_off:src_off + n]
        f = fill.reshape(1, 1, C).expand(B, T - n, C)
        return torch.cat([f, s3], 1) if ff else torch.cat([s3, f], 1)
    layers.assemble_rows = _asm
if "3" in mode:
    layers.layer_norm_rows = lambda x, w, b, eps, j0, cnt: layers.layer_norm(x[:, j0:j0 + cnt], w, b, eps)
if "4" in mode:
    ops.gather_rows = lambda src, order, j0, cnt: torch.gather(src, 1, order[:, j0:j0 + cnt, None].expand(-1, -1, src.shape[-1]))
def body():
    fp.zero_grad()
    loss = model(pts, mask=mask)
    if "c" in mode:
        model.ACT_encoder._encoded = None
    if "f" not in mode:
        loss.backward()
    return loss
s_ = torch.cuda.Stream()
s_.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s_):
    for _ in range(2):
        body()
torch.cuda.current_stream().wait_stream(s_)
torch.cuda.synchronize()
print("eager ok")
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        try:
            body()
        except Exception as e:
            print("INNER", type(e).__name__, str(e)[:300])
            traceback.print_exc()
    print("capture ok", mode)
except Exception as e:
    print("FAILED", mode, type(e).__name__, str(e)[:200])
