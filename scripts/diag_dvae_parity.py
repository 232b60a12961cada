"""Diagnostic: the Stage-I parity test body with every error printed, under switchable variants (argv[1]):
   base | atenbn (decoder BatchNorm+ReLU on ATen, f32 between layers) ."""
import sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from act_b200 import dvae, layers, ops
from act_b200.models import Cfg
from oracle import ref_model

variant = sys.argv[1] if len(sys.argv) > 1 else "base"


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


if variant == "atenbn":
    def bn_relu(x, bn, training=True):
        y = F.relu(F.batch_norm(x.float(), bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps))
        return y
    layers.bn_relu = bn_relu
    _lin = layers.linear
    layers.linear = lambda x, w, b=None, gelu=False, out_act=False: _lin(x, w, b, gelu, False)

g = np.load("tests/golden/dvae_step.npz")
with ops.precision("fp32x3"):
    cfg = Cfg(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256, tokens_dims=256,
              decoder_dims=256)
    model = ref_model.fill_params(dvae.DiscreteVAE(cfg), seed=8).cuda().train()
    pts = torch.from_numpy(g["pts"]).cuda()
    gumbel = torch.from_numpy(np.random.default_rng(41).gumbel(size=(2, 64, 8192)).astype(np.float32)).cuda()
    ret = model(pts, temperature=1.0, hard=False, gumbel=gumbel)
    l1, l2 = model.get_loss(ret, pts)
    (l1 + 0.05 * l2).backward()
    whole_coarse, whole_fine, coarse, fine, nb, logits = ret
    errs = {"logits": rel(logits[:, ::8, ::64], g["logits_sample"]), "coarse": rel(coarse, g["coarse"]),
            "fine": rel(fine, g["fine"]), "whole_fine": rel(whole_fine, g["whole_fine"]),
            "loss_recon": abs(l1.item() - float(g["loss_recon"])) / abs(float(g["loss_recon"])),
            "loss_klv": abs(l2.item() - float(g["loss_klv"])) / abs(float(g["loss_klv"]))}
    print(variant, "features", {k: f"{v:.2e}" for k, v in errs.items()})
    params = dict(model.named_parameters())
    norms = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    for k, w in norms.items():
        n = params[k].grad.norm().item()
        full = rel(params[k].grad, g["grad/" + k]) if ("grad/" + k) in g.files else float("nan")
        print(f"  {k:45s} norm_ref {w:10.3e} norm_rel {abs(n - w) / max(w, 1e-30):9.2e}  full_rel {full:9.2e}")
