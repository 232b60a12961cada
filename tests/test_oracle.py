"""CPU tests: the oracle (oracle/cpu_ref.c + oracle/ref_model.py) against the golden fixtures made from
the real reference (oracle/make_golden.py) and against independent restatements of the reference's
in-repo pure-torch FPS / kNN (utils/pc_utils.py:49-69, models/dvae.py:120-152)."""
import os

import numpy as np
import pytest
import torch

from oracle import cpu_ref, ref_model

CLOUDS = ["shapenet", "dup", "lattice", "near_origin", "n1000", "n600", "identical"]


@pytest.mark.parametrize("name", CLOUDS)
def test_group_oracle_matches_golden(golden, name):
    g = golden("group.npz")
    xyz = g[name + "/xyz"]
    nb, center, idx, fps_idx = cpu_ref.group(xyz, 64, 32)
    assert np.array_equal(fps_idx, g[name + "/fps_idx"])
    assert np.array_equal(idx, g[name + "/knn_idx"].astype(np.int64))
    assert np.array_equal(center, g[name + "/center"])
    assert np.array_equal(nb, g[name + "/neighborhood"])


def test_fps_matches_pure_torch_semantics(golden):
    """utils/pc_utils.py:49-69 recurrence (min-distance update + argmax), start index forced to 0.
    On tie-free data with no point near the origin the index sequence must be identical."""
    xyz = golden("group.npz")["shapenet/xyz"]
    got = cpu_ref.fps(xyz, 64)
    for b in range(xyz.shape[0]):
        p = xyz[b].astype(np.float32)
        assert (np.sum(p * p, -1) > 1e-3).all()
        dist = np.full(p.shape[0], 1e10, np.float32)
        far, seq = 0, []
        for _ in range(64):
            seq.append(far)
            d = ((p - p[far]) ** 2).sum(-1).astype(np.float32)
            dist = np.minimum(dist, d)
            far = int(dist.argmax())
        assert (np.array(seq) == got[b]).mean() > 0.97      # fp32 summation order differs in the last ulp


def test_knn_matches_knn_point_set(golden):
    """models/dvae.py:120-152 knn_point (expanded-form distance + unsorted topk): same neighbour set."""
    g = golden("group.npz")
    xyz, center = torch.from_numpy(g["shapenet/xyz"]), torch.from_numpy(g["shapenet/center"])
    d = -2 * center @ xyz.transpose(1, 2) + (center ** 2).sum(-1)[..., None] + (xyz ** 2).sum(-1)[:, None]
    ref = d.topk(32, dim=-1, largest=False)[1].sort(-1)[0].numpy()
    dist, idx = cpu_ref.knn(g["shapenet/xyz"], g["shapenet/center"], 32)
    assert (np.sort(idx, -1) == ref).mean() > 0.999
    assert (np.diff(dist, axis=-1) >= 0).all()
    assert (idx[:, :, 0] == g["shapenet/fps_idx"]).all()       # nearest neighbour of a centre is itself


def test_knn_edge_cases():
    rng = np.random.default_rng(0)
    ref = rng.standard_normal((1, 40, 3)).astype(np.float32)
    d, i = cpu_ref.knn(ref, ref[:, :5], 40)                    # k == N
    assert sorted(i[0, 0].tolist()) == list(range(40))
    same = np.zeros((1, 16, 3), np.float32)                    # all distances equal -> index order
    d, i = cpu_ref.knn(same, same[:, :2], 4)
    assert i[0, 0].tolist() == [0, 1, 2, 3]


def test_chamfer_oracle_bruteforce():
    rng = np.random.default_rng(1)
    for n, m in [(8, 32), (32, 32), (600, 1030), (1, 1)]:
        a = rng.standard_normal((3, n, 3)).astype(np.float32)
        b = rng.standard_normal((3, m, 3)).astype(np.float32)
        d1, d2, i1, i2 = cpu_ref.chamfer_forward(a, b)
        full = ((a[:, :, None, :].astype(np.float64) - b[:, None, :, :]) ** 2).sum(-1)
        assert np.array_equal(i1, full.argmin(2)) and np.array_equal(i2, full.argmin(1))
        np.testing.assert_allclose(d1, full.min(2), rtol=1e-5, atol=1e-6)
        g1 = rng.standard_normal(d1.shape).astype(np.float32)
        g2 = rng.standard_normal(d2.shape).astype(np.float32)
        gx1, gx2 = cpu_ref.chamfer_backward(a, b, i1, i2, g1, g2)
        ta, tb = torch.tensor(a, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True)
        dd = ((ta[:, :, None] - tb[:, None]) ** 2).sum(-1)
        ((dd.min(2)[0] * torch.tensor(g1)).sum() + (dd.min(1)[0] * torch.tensor(g2)).sum()).backward()
        np.testing.assert_allclose(gx1, ta.grad.numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(gx2, tb.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_chamfer_ties_lowest_index():
    a = np.zeros((1, 4, 3), np.float32)
    b = np.zeros((1, 700, 3), np.float32)                      # spans two 512-tiles, all tied
    _, _, i1, i2 = cpu_ref.chamfer_forward(a, b)
    assert (i1 == 0).all() and (i2 == 0).all()


def test_block12_cfg1_restatement_matches_reference_golden(golden):
    """BASELINE config 1 through the oracle's Block restatement == utils/transformer_layers.Block x12."""
    g = golden("block12_cfg1.npz")
    blocks = torch.nn.ModuleList([ref_model.Block(384, 6) for _ in range(12)]).eval()
    ref_model.fill_params(blocks, seed=1)
    y = torch.from_numpy(g["x"])
    with torch.no_grad():
        for b in blocks:
            y = b(y)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-4, atol=1e-5)


def test_encoder_restatement_matches_reference_golden(golden):
    g, grp = golden("encoder.npz"), golden("group.npz")
    nb = torch.from_numpy(grp["shapenet/neighborhood"][:2]).requires_grad_(True)
    enc = ref_model.fill_params(ref_model.Encoder(384), seed=2).train()
    out = enc(nb)
    (out * torch.from_numpy(g["wout"])).sum().backward()
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(nb.grad.numpy(), g["grad_in"], rtol=1e-3, atol=1e-5)
    for k, p in enc.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), g["grad/" + k], rtol=2e-3, atol=2e-4)
    for k, b in enc.named_buffers():
        np.testing.assert_allclose(b.numpy(), g["buf/" + k], rtol=1e-5, atol=1e-6)


def test_student_step_restatement_matches_reference_golden(golden):
    g = golden("student_step.npz")
    model = ref_model.fill_params(ref_model.ACTPointDistillationStudent(mask_ratio=0.6), seed=4).train()
    loss = model(torch.from_numpy(g["pts"]), torch.from_numpy(g["teacher"]), torch.from_numpy(g["mask"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    norms = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    n = 0
    for k, p in model.named_parameters():
        if p.grad is None:
            assert k not in norms
            continue
        assert abs(p.grad.norm().item() - norms[k]) <= 2e-3 * norms[k] + 1e-9, k
        if "grad/" + k in g.files:
            np.testing.assert_allclose(p.grad.numpy(), g["grad/" + k], rtol=2e-3, atol=1e-7)
        n += 1
    assert n == len(norms) == 183


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="authoring container only")
def test_state_dict_keys_match_real_reference():
    """The restatement's (and hence the product's) state_dict keys/shapes == models/act.py's student."""
    from oracle import shims
    shims.install()
    import models.act as act
    from models.dvae import Encoder
    ref_keys = {k: tuple(v.shape) for k, v in Encoder(384).state_dict().items()}
    mine = {k: tuple(v.shape) for k, v in ref_model.Encoder(384).state_dict().items()}
    assert ref_keys == mine
    te = act.TransformerEncoder(embed_dim=384, depth=2, num_heads=6, drop_path_rate=[0.0, 0.0])
    assert {k: tuple(v.shape) for k, v in te.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in ref_model.TransformerEncoder(384, 2, 6).state_dict().items()}


def test_product_stage1_and_finetune_modules_have_the_oracle_state_dict():
    """act_b200.dvae.DiscreteVAE / act_b200.models.PointTransformer (constructed on the CPU: parameters only) carry the
    same state_dict keys and shapes as the oracle restatements, which tests/test_oracle_dvae.py and
    tests/test_point_transformer.py pin to the unmodified reference classes."""
    from act_b200 import dvae, models
    from oracle import ref_dvae, ref_model

    def shapes(m):
        return {k: tuple(v.shape) for k, v in m.state_dict().items()}

    cfg = models.Cfg(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256, tokens_dims=256,
                     decoder_dims=256)
    assert shapes(dvae.DiscreteVAE(cfg)) == shapes(ref_dvae.DiscreteVAE())
    assert models.MODELS["DiscreteVAE"] is dvae.DiscreteVAE
    for tt in ("full", "linear", "side"):
        pc = models.Cfg(NAME="PointTransformer", embed_dim=384, depth=12, drop_path_rate=0.1, cls_dim=40, num_heads=6,
                        group_size=32, num_group=64, encoder_dims=384, transfer_type=tt)
        assert shapes(models.PointTransformer(pc)) == shapes(ref_model.PointTransformer(transfer_type=tt))


def test_stage1_schedules_product_equals_oracle():
    from act_b200 import dvae
    from oracle import ref_dvae
    for n in (0, 1, 9999, 10000, 10001, 55000, 100000, 100001, 110000, 110001, 10 ** 6):
        assert dvae.get_temp(n) == ref_dvae.temperature_schedule(n)
        assert dvae.get_kld_weight(n) == ref_dvae.kld_weight_schedule(n)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="authoring container only")
def test_block_mask_restatement_matches_reference():
    """oracle.ref_model.mask_center_block against the unmodified VisableOnlyMaskTransformer._mask_center_block
    (act.py:215-243) on the same Python `random` stream."""
    import random
    import types
    from oracle import ref_model, shims
    shims.install()
    import models.act as act
    center = torch.randn(5, 64, 3)
    stub = types.SimpleNamespace(mask_ratio=0.6)
    random.seed(3)
    want = act.VisableOnlyMaskTransformer._mask_center_block(stub, center)
    random.seed(3)
    got = ref_model.mask_center_block(center, 0.6)
    assert torch.equal(want, got) and int(got.sum()) == 5 * int(0.6 * 64)
