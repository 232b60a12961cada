"""GPU: the UNMODIFIED reference modules running over dropin/ (pointnet2_ops, knn_cuda, chamfer backed by libact_b200.so)
-- the "drop-in native ops" integration path of INTEGRATION.md section 1, exercised end to end.

Needs a reference tree (read-only, never part of this repo): $ACT_REFERENCE_TREE, /root/reference (authoring container), or
<repo>/_scratch_ref (an uncommitted, git-ignored scratch copy pushed to the GPU box for one run:
    mkdir -p _scratch_ref && cp -r /root/reference/{models,utils,extensions} _scratch_ref/ && gpurun ... ; rm -rf _scratch_ref).
Skipped when none is reachable.  Runs in a SUBPROCESS: the reference's import names (`models`, `utils`, `chamfer`,
`knn_cuda`, ...) and the stand-ins for its non-native dependencies must not leak into the other tests' interpreter."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tree():
    for p in (os.environ.get("ACT_REFERENCE_TREE"), "/root/reference", os.path.join(ROOT, "_scratch_ref")):
        if p and os.path.isfile(os.path.join(p, "models", "dvae.py")):
            return p
    return None


SCRIPT = r'''
import json, os, sys
import numpy as np, torch
ROOT, TREE = sys.argv[1], sys.argv[2]
sys.path.insert(0, ROOT)
# the reference's Conv1d / Linear run on cuDNN / cuBLAS here: hold them to plain fp32 (cuDNN convolutions default to TF32,
# which alone moves these outputs by ~5e-4 -- the "TF32-grade reference numerics" of SURVEY.md 8a)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import shims, ref_model
shims.install(native="dropin", reference_root=TREE)
import knn_cuda, chamfer, pointnet2_ops.pointnet2_utils as p2u           # must resolve to dropin/
assert os.path.realpath(knn_cuda.__file__).startswith(os.path.realpath(os.path.join(ROOT, "dropin")))
assert os.path.realpath(chamfer.__file__).startswith(os.path.realpath(os.path.join(ROOT, "dropin")))
import models.dvae as dvae                                               # the reference's own file, unmodified
from extensions.chamfer_dist import ChamferDistanceL1, ChamferDistanceL2  # the reference's autograd wrappers over `chamfer`
G = os.path.join(ROOT, "tests", "golden")
out = {}
rel = lambda a, b: float((torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()).norm() / (torch.as_tensor(b).double().norm() + 1e-30))

# Group (dvae.py:154-183): misc.fps -> furthest_point_sample + gather_operation, KNN(k=32, transpose_mode=True)
grp = np.load(os.path.join(G, "group.npz"))
pts = torch.from_numpy(grp["shapenet/xyz"]).cuda()
nb, center = dvae.Group(num_group=64, group_size=32)(pts)
out["group_neighborhood_equal"] = bool(np.array_equal(nb.cpu().numpy(), grp["shapenet/neighborhood"]))
out["group_center_equal"] = bool(np.array_equal(center.cpu().numpy(), grp["shapenet/center"]))

# Encoder (dvae.py:185-215): pure PyTorch in the reference; here it only has to run on the drop-in's output
g = np.load(os.path.join(G, "encoder.npz"))
enc = ref_model.fill_params(dvae.Encoder(384), seed=2).cuda().train()
o = enc(nb[:2])
out["encoder_rel"] = rel(o, g["out"])

# DGCNN (dvae.py:26-117): KNN(k=4, transpose_mode=False) through the drop-in, [B,3,N] layout
t = np.load(os.path.join(G, "teacher.npz"))
cfg = shims.easydict(dict(group_size=32, num_group=64, encoder_dims=384, tokens_dims=384, decoder_dims=384, num_tokens=8192))
dg = ref_model.fill_params(dvae.DGCNN(encoder_channel=384, output_channel=512), seed=12).cuda()
x = torch.from_numpy(np.random.default_rng(2).standard_normal((2, 64, 384)).astype(np.float32)).cuda()
from oracle import ref_teacher
want = ref_model.fill_params(ref_teacher.DGCNN(384, 512), seed=12)(x.cpu(), center[:2].cpu())
out["dgcnn_rel"] = rel(dg(x, center[:2]), want.detach())

# ChamferDistanceL1 / L2 (extensions/chamfer_dist/__init__.py) forward + backward through dropin/chamfer.py
from oracle import cpu_ref
rng = np.random.default_rng(5)
a = torch.from_numpy(rng.standard_normal((64, 32, 3)).astype(np.float32)).cuda().requires_grad_(True)
b = torch.from_numpy(rng.standard_normal((64, 48, 3)).astype(np.float32)).cuda().requires_grad_(True)
l2 = ChamferDistanceL2()(a, b)
l1 = ChamferDistanceL1()(a, b)
(l1 + l2).backward()
d1, d2, i1, i2 = cpu_ref.chamfer_forward(a.detach().cpu().numpy(), b.detach().cpu().numpy())
out["chamfer_l2_rel"] = abs(l2.item() - float(d1.mean() + d2.mean())) / float(d1.mean() + d2.mean())
out["chamfer_l1_rel"] = abs(l1.item() - (np.sqrt(d1).mean() + np.sqrt(d2).mean()) / 2) / ((np.sqrt(d1).mean() + np.sqrt(d2).mean()) / 2)
out["chamfer_grad_finite"] = bool(torch.isfinite(a.grad).all() and torch.isfinite(b.grad).all() and a.grad.abs().sum() > 0)
print("RESULT " + json.dumps({k: (v if isinstance(v, bool) else float(v)) for k, v in out.items()}))
'''


def test_unmodified_reference_modules_over_dropin():
    tree = _tree()
    if tree is None:
        pytest.skip("no reference tree reachable (ACT_REFERENCE_TREE, /root/reference, <repo>/_scratch_ref)")
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, tree], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[7:])
    assert out["group_neighborhood_equal"] and out["group_center_equal"], out      # bit-exact tokenizer through the drop-ins
    assert out["encoder_rel"] < 2e-4, out                                          # the reference's fp32 PyTorch Encoder
    assert out["dgcnn_rel"] < 2e-4, out                                            # same neighbours as the oracle's kNN
    assert out["chamfer_l2_rel"] < 1e-5 and out["chamfer_l1_rel"] < 1e-5 and out["chamfer_grad_finite"], out
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dropin_reference.json"), "w") as f:
        json.dump(out, f, indent=1)
