"""CPU: the oracle restatement of the frozen teacher's feature path (oracle/ref_teacher.py) against the golden
fixture made from the unmodified reference class, and -- in the authoring container -- against the reference itself."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_model, ref_teacher


def _inputs(golden):
    g = golden("group.npz")
    return torch.from_numpy(g["shapenet/neighborhood"][:2]), torch.from_numpy(g["shapenet/center"][:2])


def _noise(B=2, G=64):
    rng = np.random.default_rng(31)                      # == oracle.make_golden.teacher_noise
    gumbel = torch.from_numpy(rng.gumbel(size=(B, G, 8192)).astype(np.float32))
    keeps = [torch.from_numpy((rng.random((B, 64, 768)) >= 0.1).astype(np.float32)) for _ in range(12)]
    return gumbel, keeps


def test_teacher_restatement_matches_reference_golden(golden):
    g = golden("teacher.npz")
    nb, center = _inputs(golden)
    torch.set_num_threads(8)
    model = ref_model.fill_params(ref_teacher.TeacherFeatures(), seed=6).train()
    gumbel, keeps = _noise()
    with torch.no_grad():
        logits = model.dgcnn_1(model.encoder(nb), center)
        model.encoder.first_conv[1].reset_running_stats()
        model.encoder.second_conv[1].reset_running_stats()
        feat = model.forward_tokenizer_features(nb, center, gumbel=gumbel, keeps=keeps)
    np.testing.assert_allclose(logits[:, ::8, ::64].numpy(), g["logits_sample"], rtol=2e-4, atol=2e-4)
    assert ((logits + gumbel).argmax(-1).numpy() == g["labels"]).all()
    np.testing.assert_allclose(feat.numpy(), g["feature"], rtol=1e-3, atol=1e-4)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="authoring container only")
def test_teacher_state_dict_keys_match_real_reference():
    from oracle import shims
    shims.install()
    import models.dvae as dvae
    cfg = shims.easydict(dict(group_size=32, num_group=64, encoder_dims=384, tokens_dims=384, decoder_dims=384,
                              num_tokens=8192, visual_embed_type="vit_base_patch16_384", visual_embed_dim=768,
                              freeze_visual_embed=True, num_prompt_token=64, use_deep_prompt=True))
    ref = dvae.ACTPromptedDiscreteVAEwithVIT(cfg)
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref_teacher.TeacherFeatures().state_dict().items()}
    assert a == b, (set(a) ^ set(b))
