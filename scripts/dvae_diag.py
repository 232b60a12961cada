"""Prints every parity figure of the Stage-I dVAE step against tests/golden/dvae_step.npz (diagnostic for tolerances)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from act_b200 import dvae
from act_b200.models import Cfg
from oracle import ref_model
if os.environ.get("DVAE_FP32_LINEAR"):      # structure check: fp32 library Linear instead of the bf16 tcgen05 GEMM
    from act_b200 import layers
    layers.linear = lambda x, w, b=None, gelu=False: torch.nn.functional.linear(x, w, b)
g = np.load("tests/golden/dvae_step.npz")
cfg = Cfg(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256, tokens_dims=256, decoder_dims=256)
model = ref_model.fill_params(dvae.DiscreteVAE(cfg), seed=8).cuda().train()
pts = torch.from_numpy(g["pts"]).cuda()
gum = torch.from_numpy(np.random.default_rng(41).gumbel(size=(2, 64, 8192)).astype(np.float32)).cuda()
ret = model(pts, temperature=1.0, hard=False, gumbel=gum)
l1, l2 = model.get_loss(ret, pts)
(l1 + 0.05 * l2).backward()
rel = lambda a, b: ((torch.as_tensor(a).float().cpu() - torch.as_tensor(b).float()).norm() / torch.as_tensor(b).float().norm()).item()
out = {"logits": rel(ret[5][:, ::8, ::64], g["logits_sample"]), "coarse": rel(ret[2], g["coarse"]), "fine": rel(ret[3], g["fine"]),
       "l1": (l1.item(), float(g["loss_recon"])), "l2": (l2.item(), float(g["loss_klv"]))}
P = dict(model.named_parameters())
out["norms"] = {k: (P[k].grad.norm().item(), w) for k, w in zip(g["grad_names"].tolist(), g["grad_norms"].tolist())}
out["full"] = {k: rel(P[k[5:]].grad, g[k]) for k in g.files if k.startswith("grad/") and k != "grad/codebook_rows"}
out["codebook"] = rel(model.codebook.grad[::512], g["grad/codebook_rows"])
print(json.dumps(out, indent=1))
