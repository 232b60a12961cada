"""CPU: the oracle restatement of the Stage-I dVAE training step (oracle/ref_dvae.py, SURVEY row f2 / BASELINE
config 3) against the golden fixture made from the unmodified reference `DiscreteVAE`, and -- in the authoring
container -- against the reference itself."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import ref_dvae, ref_model

KLD_WEIGHT = 0.05                                       # == oracle.make_golden.DVAE_KLD_WEIGHT


def _noise(B=2, G=64):
    return torch.from_numpy(np.random.default_rng(41).gumbel(size=(B, G, 8192)).astype(np.float32))   # == dvae_noise


def _step(golden):
    g = golden("dvae_step.npz")
    torch.set_num_threads(8)
    model = ref_model.fill_params(ref_dvae.DiscreteVAE(), seed=8).train()
    pts = torch.from_numpy(g["pts"])
    ret = model(pts, temperature=1.0, hard=False, gumbel=_noise())
    l1, l2 = model.get_loss(ret, pts)
    (l1 + KLD_WEIGHT * l2).backward()
    return g, model, ret, l1, l2


def test_dvae_restatement_matches_reference_golden(golden):
    g, model, ret, l1, l2 = _step(golden)
    _, whole_fine, coarse, fine, _, logits = ret
    np.testing.assert_allclose(logits.detach()[:, ::8, ::64].numpy(), g["logits_sample"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(coarse.detach().numpy(), g["coarse"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(fine.detach().numpy(), g["fine"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(whole_fine.numpy(), g["whole_fine"], rtol=1e-3, atol=1e-4)
    assert abs(l1.item() - g["loss_recon"]) <= 1e-4 * abs(g["loss_recon"])
    assert abs(l2.item() - g["loss_klv"]) <= 1e-3 * abs(g["loss_klv"])
    grads = dict(model.named_parameters())
    norms = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    assert set(norms) == {k for k, p in grads.items() if p.grad is not None}
    for k, want in norms.items():
        got = grads[k].grad.norm().item()
        assert abs(got - want) <= 2e-3 * want + 1e-7, (k, got, want)
    for k in g.files:
        if k.startswith("grad/") and k != "grad/codebook_rows":
            np.testing.assert_allclose(grads[k[5:]].grad.numpy(), g[k], rtol=5e-3, atol=1e-5 + 2e-3 * np.abs(g[k]).max())
    np.testing.assert_allclose(model.codebook.grad[::512].numpy(), g["grad/codebook_rows"], rtol=5e-3,
                               atol=2e-3 * np.abs(g["grad/codebook_rows"]).max())
    for k, b in model.named_buffers():
        if "running" in k:
            np.testing.assert_allclose(b.numpy(), g["buf/" + k], rtol=1e-4, atol=1e-5)


def test_dvae_restatement_smooth_loss_gradients(golden):
    """The step differentiated through the smooth surrogate loss (make_golden.gen_dvae_step_smooth): the fixture the GPU
    parity test holds the whole Stage-I backward to, free of Chamfer-L1's arg-min discontinuity."""
    g, gs = golden("dvae_step.npz"), golden("dvae_step_smooth.npz")
    torch.set_num_threads(8)
    model = ref_model.fill_params(ref_dvae.DiscreteVAE(), seed=8).train()
    pts = torch.from_numpy(g["pts"])
    ret = model(pts, temperature=1.0, hard=False, gumbel=_noise())
    _, _, coarse, fine, _, _ = ret
    _, l2 = model.get_loss(ret, pts)
    rng = np.random.default_rng(43)
    rc = torch.from_numpy(rng.standard_normal(tuple(coarse.shape)).astype(np.float32)) / float(np.prod(coarse.shape[:-1]))
    rf = torch.from_numpy(rng.standard_normal(tuple(fine.shape)).astype(np.float32)) / float(np.prod(fine.shape[:-1]))
    loss = (coarse * rc).sum() + (fine * rf).sum() + KLD_WEIGHT * l2
    loss.backward()
    assert abs(loss.item() - float(gs["loss"])) <= 1e-4 * abs(float(gs["loss"])) + 1e-8
    grads = dict(model.named_parameters())
    floor = 1e-5 * gs["grad_norms"].max()          # biases in front of a BatchNorm: analytically zero gradient, noise only
    for k, want in zip(gs["grad_names"].tolist(), gs["grad_norms"].tolist()):
        got = grads[k].grad.norm().item()
        assert want < floor or abs(got - want) <= 1e-3 * want, (k, got, want)
    for k in gs.files:
        if k.startswith("grad/") and k != "grad/codebook_rows":
            np.testing.assert_allclose(grads[k[5:]].grad.numpy(), gs[k], rtol=2e-3, atol=1e-3 * np.abs(gs[k]).max())


def test_schedules():
    """tools/runner_autoencoder.py:18-53 with cfgs/autoencoder/pointbert_dvae.yaml:27-38."""
    assert ref_dvae.temperature_schedule(0) == 1.0
    assert abs(ref_dvae.temperature_schedule(50000) - (0.0625 + 0.9375 / 2)) < 1e-12
    assert ref_dvae.temperature_schedule(100001) == 0.0625
    assert ref_dvae.kld_weight_schedule(9999) == 0.0
    assert ref_dvae.kld_weight_schedule(10000) == 0.0
    assert abs(ref_dvae.kld_weight_schedule(60000) - 0.05) < 1e-12
    assert ref_dvae.kld_weight_schedule(110001) == 0.1
    assert math.isclose(ref_dvae.kld_weight_schedule(110000), 0.1)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="authoring container only")
def test_dvae_state_dict_keys_match_real_reference():
    from oracle import shims
    shims.install()
    import models.dvae as dvae
    cfg = shims.easydict(dict(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256,
                              tokens_dims=256, decoder_dims=256))
    ref = dvae.DiscreteVAE(cfg)
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref_dvae.DiscreteVAE().state_dict().items()}
    assert a == b, (set(a) ^ set(b))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="authoring container only")
def test_dvae_restatement_eval_hard_path_matches_real_reference():
    """runner_autoencoder.py:240 (`base_model(inp=points, hard=True, eval=True)`): eval-mode BatchNorm, straight-through
    one-hot.  The oracle restatement against the unmodified reference class on other clouds and another temperature."""
    from oracle import shims
    shims.install()
    import models.dvae as dvae
    cfg = shims.easydict(dict(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256,
                              tokens_dims=256, decoder_dims=256))
    torch.set_num_threads(8)
    ref = ref_model.fill_params(dvae.DiscreteVAE(cfg), seed=3).eval()
    mine = ref_model.fill_params(ref_dvae.DiscreteVAE(), seed=3).eval()
    pts = ref_model.synthetic_clouds(2, 1024, seed=77)
    gum = torch.from_numpy(np.random.default_rng(5).gumbel(size=(2, 64, 8192)).astype(np.float32))
    orig = dvae.F.gumbel_softmax
    dvae.F.gumbel_softmax = lambda logits, tau=1.0, hard=False, dim=-1: ref_dvae.gumbel_softmax_with_noise(logits, gum, tau, hard)
    try:
        with torch.no_grad():
            want = ref(inp=pts, temperature=0.5, hard=True, eval=True)
    finally:
        dvae.F.gumbel_softmax = orig
    with torch.no_grad():
        got = mine(pts, temperature=0.5, hard=True, gumbel=gum, eval=True)
    for a, b in zip(got, want):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-3, atol=1e-4)
