"""Generate tests/golden/*.npz from the UNMODIFIED reference (imported through oracle/shims.py).
TEST INFRASTRUCTURE ONLY.  Run in the authoring container:  python -m oracle.make_golden

The reference ships no golden vectors, so these are outputs of the reference's
own Python modules run here on CPU fp32 -- models/dvae.py (Group, Encoder), models/act.py
(TransformerEncoder, ACT_PointDistillation), utils/transformer_layers.py (Block, BASELINE config 1)
-- on seeded inputs with the deterministic weights of oracle.ref_model.fill_params.  The native ops
under Group are the C oracle (the upstream CUDA packages cannot run here).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")

from . import shims, cpu_ref                      # noqa: E402
from .ref_model import fill_params, synthetic_clouds  # noqa: E402


def adversarial_clouds():
    """SURVEY.md 8(d) parity set: duplicates, lattice (many equal distances), >=8 points with
    |p|^2 <= 1e-3, N not a multiple of 512, all-identical cloud.  dict name -> [B,N,3] f32."""
    rng = np.random.default_rng(7)
    out = {}
    base = synthetic_clouds(2, 1024, seed=11).numpy()
    dup = base.copy()
    dup[:, 512:768] = dup[:, 0:256]                       # exact duplicates
    out["dup"] = dup
    g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(16), indexing="ij"), -1).reshape(-1, 3)
    lat = (g.astype(np.float32) - np.array([3.5, 3.5, 7.5], np.float32)) * 0.125
    out["lattice"] = np.stack([lat, lat[rng.permutation(1024)]])
    org = base.copy()
    org[:, 5:17] = (rng.standard_normal((2, 12, 3)) * 0.01).astype(np.float32)   # |p|^2 <= 1e-3
    out["near_origin"] = org
    out["n1000"] = synthetic_clouds(2, 1000, seed=12).numpy()
    out["n600"] = synthetic_clouds(3, 600, seed=13).numpy()
    out["identical"] = np.full((1, 1024, 3), 0.25, np.float32)
    return out


def gen_group():
    from models.dvae import Group, knn_point
    res = {}
    clouds = {"shapenet": synthetic_clouds(4, 1024).numpy()}
    clouds.update(adversarial_clouds())
    for name, xyz in clouds.items():
        grp = Group(64, 32)
        x = torch.from_numpy(xyz)
        nb, center = grp(x)                                  # reference Group.forward over the C oracle
        fps_idx = cpu_ref.fps(xyz, 64)
        _, idx = cpu_ref.knn(xyz, center.numpy(), 32)
        res[name + "/xyz"] = xyz
        res[name + "/fps_idx"] = fps_idx
        res[name + "/knn_idx"] = idx.astype(np.int32)
        res[name + "/center"] = center.numpy()
        res[name + "/neighborhood"] = nb.numpy()
        if name == "shapenet":                               # tie-free: same neighbour SET as the in-repo knn_point
            s = knn_point(32, x, center).sort(-1)[0].numpy()
            assert (np.sort(idx, -1) == s).mean() > 0.999, "oracle kNN disagrees with reference knn_point"
    np.savez_compressed(os.path.join(GOLD, "group.npz"), **res)
    print("group.npz", {k: v.shape for k, v in res.items() if k.startswith("shapenet")})


def gen_block_cfg1():
    """BASELINE config 1: utils/transformer_layers.Block x12, d=384, 64 tokens, batch 2, CPU, eval."""
    from utils.transformer_layers import Block
    blocks = torch.nn.ModuleList([Block(384, 6) for _ in range(12)]).eval()
    fill_params(blocks, seed=1)
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((2, 64, 384)).astype(np.float32))
    with torch.no_grad():
        y = x
        for b in blocks:
            y = b(y)
    np.savez_compressed(os.path.join(GOLD, "block12_cfg1.npz"), x=x.numpy(), y=y.numpy())
    print("block12_cfg1.npz", y.abs().mean().item())


def gen_encoder():
    from models.dvae import Encoder
    g = np.load(os.path.join(GOLD, "group.npz"))
    nb = torch.from_numpy(g["shapenet/neighborhood"][:2])
    enc = fill_params(Encoder(384), seed=2).train()
    nb.requires_grad_(True)
    out = enc(nb)
    w = torch.from_numpy(np.random.default_rng(3).standard_normal(out.shape).astype(np.float32))
    (out * w).sum().backward()
    res = {"out": out.detach().numpy(), "wout": w.numpy(), "grad_in": nb.grad.numpy()}
    for k, p in enc.named_parameters():
        res["grad/" + k] = p.grad.numpy()
    for k, b in enc.named_buffers():
        res["buf/" + k] = b.numpy()
    enc.eval()
    with torch.no_grad():
        res["out_eval"] = enc(nb.detach()).numpy()
    np.savez_compressed(os.path.join(GOLD, "encoder.npz"), **res)
    print("encoder.npz", out.abs().mean().item())


def gen_student_step():
    """ACT_PointDistillation.forward/backward of the real reference (act.py:1203-1258), B=4, mask 0.6,
    drop_path 0, with the frozen teacher replaced by a stub returning seeded features."""
    import models.act as act
    B, G = 4, 64
    cfg = shims.easydict(dict(
        NAME="ACT_PointDistillation", loss="cosine",
        transformer_config=dict(mask_ratio=0.6, mask_type="rand", proj="linear", embed_dim=384, encoder_dims=384,
                                depth=12, drop_path_rate=0.0, cls_dim=512, replace_pob=0.0, num_heads=6,
                                decoder_depth=2, decoder_num_heads=6, return_all_tokens=False, cls_loss=False,
                                register_shallow_hook=9),
        dvae_config=dict(num_group=G, group_size=32, encoder_dims=384, num_tokens=8192, tokens_dims=384,
                         decoder_dims=384, ckpt="")))
    teacher = torch.from_numpy(np.random.default_rng(5).standard_normal((B, G, 384)).astype(np.float32))

    class StubTokenizer(torch.nn.Module):
        def forward_tokenizer_features(self, neighborhood, center, return_global=True):
            return teacher

    def build_tokenizer(self, cfg_):
        self.dvae_tokenizer = StubTokenizer()

    act.ACT_PointDistillation.build_tokenizer = build_tokenizer
    model = act.ACT_PointDistillation(cfg)
    fill_params(model, seed=4)
    model.train()
    pts = synthetic_clouds(B, 1024, seed=21)
    np.random.seed(123)
    captured = {}
    orig = act.VisableOnlyMaskTransformer._mask_center_rand

    def capture(self, center, noaug=False):
        m = orig(self, center, noaug)
        captured["mask"] = m.clone()
        return m

    act.VisableOnlyMaskTransformer._mask_center_rand = capture
    loss = model(pts)
    loss.backward()
    res = {"pts": pts.numpy(), "teacher": teacher.numpy(), "mask": captured["mask"].numpy(),
           "loss": np.float32(loss.item())}
    keep_full = ("proj_head.bias", "mask_token", "ACT_encoder.cls_token", "ACT_encoder.encoder.first_conv.0.weight",
                 "ACT_encoder.blocks.blocks.0.norm1.weight", "ACT_decoder.norm.bias",
                 "ACT_encoder.pos_embed.0.weight", "ACT_encoder.blocks.blocks.11.attn.proj.bias")
    names, norms = [], []
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        names.append(k)
        norms.append(p.grad.norm().item())
        if k in keep_full:
            res["grad/" + k] = p.grad.numpy()
    res["grad_names"] = np.array(names)
    res["grad_norms"] = np.array(norms, np.float64)
    for k, b in model.named_buffers():
        if "running" in k:
            res["buf/" + k] = b.numpy()
    np.savez_compressed(os.path.join(GOLD, "student_step.npz"), **res)
    print("student_step.npz loss", loss.item(), "n grads", len(names))


def teacher_noise(B, G, num_tokens, depth, P, seed=31):
    """Fixed gumbel noise and prompt-dropout keep masks shared by the golden run, the oracle and the GPU tests."""
    rng = np.random.default_rng(seed)
    gumbel = torch.from_numpy(rng.gumbel(size=(B, G, num_tokens)).astype(np.float32))
    keeps = [torch.from_numpy((rng.random((B, P, 768)) >= 0.1).astype(np.float32)) for _ in range(depth)]
    return gumbel, keeps


def run_reference_teacher(nb, center, seed=6):
    """The UNMODIFIED ACTPromptedDiscreteVAEwithVIT.forward_tokenizer_features (dvae.py:584-592) in train mode, with
    its two RNG consumers (F.gumbel_softmax, prompt_dropout) replaced by the fixed noise of teacher_noise()."""
    import models.dvae as dvae
    cfg = shims.easydict(dict(group_size=32, num_group=64, encoder_dims=384, tokens_dims=384, decoder_dims=384,
                              num_tokens=8192, visual_embed_type="vit_base_patch16_384", visual_embed_dim=768,
                              freeze_visual_embed=True, num_prompt_token=64, use_deep_prompt=True))
    model = dvae.ACTPromptedDiscreteVAEwithVIT(cfg)
    fill_params(model, seed=seed)
    model.train()
    B, G = center.shape[:2]
    gumbel, keeps = teacher_noise(B, G, 8192, 12, 64)
    calls = {"i": 0}

    class FixedDrop(torch.nn.Module):
        def forward(self, t):
            k = keeps[calls["i"]]
            calls["i"] += 1
            return t * k / 0.9

    model.prompt_dropout = FixedDrop()
    orig = dvae.F.gumbel_softmax
    labels = {}

    def fixed_gumbel(logits, tau=1.0, hard=False, dim=-1):
        idx = (logits + gumbel).argmax(dim)
        labels["idx"], labels["logits"] = idx, logits
        return torch.nn.functional.one_hot(idx, logits.shape[dim]).to(logits.dtype)

    dvae.F.gumbel_softmax = fixed_gumbel
    try:
        with torch.no_grad():
            feat = model.forward_tokenizer_features(nb, center, return_global=True)
    finally:
        dvae.F.gumbel_softmax = orig
    return model, feat, labels


def gen_teacher():
    g = np.load(os.path.join(GOLD, "group.npz"))
    nb = torch.from_numpy(g["shapenet/neighborhood"][:2])
    center = torch.from_numpy(g["shapenet/center"][:2])
    _, feat, labels = run_reference_teacher(nb, center)
    lg = labels["logits"]
    np.savez_compressed(os.path.join(GOLD, "teacher.npz"), feature=feat.numpy(), labels=labels["idx"].numpy().astype(np.int32),
                        logits_sample=lg[:, ::8, ::64].numpy(), logits_absmean=np.float32(lg.abs().mean().item()))
    print("teacher.npz", feat.abs().mean().item(), labels["idx"][0, :6].tolist())


def dvae_noise(B, G, num_tokens, seed=41):
    """Fixed gumbel noise shared by the golden run, the oracle and the GPU tests of the Stage-I step."""
    return torch.from_numpy(np.random.default_rng(seed).gumbel(size=(B, G, num_tokens)).astype(np.float32))


DVAE_KLD_WEIGHT = 0.05      # a mid-schedule value (runner_autoencoder.py:18-41), so the KL gradient is exercised too


def run_reference_dvae(pts, temperature=1.0, seed=8):
    """One Stage-I training step of the UNMODIFIED reference DiscreteVAE (dvae.py:278-357) as
    tools/runner_autoencoder.py:137-146 drives it (forward, get_loss, loss_1 + kld_weight * loss_2, backward), with
    F.gumbel_softmax's internal noise replaced by dvae_noise()."""
    import models.dvae as dvae
    from .ref_dvae import gumbel_softmax_with_noise
    cfg = shims.easydict(dict(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256,
                              tokens_dims=256, decoder_dims=256))      # cfgs/autoencoder/pointbert_dvae.yaml:41-48
    model = dvae.DiscreteVAE(cfg)
    fill_params(model, seed=seed)
    model.train()
    B = pts.shape[0]
    gumbel = dvae_noise(B, 64, 8192)
    orig = dvae.F.gumbel_softmax
    dvae.F.gumbel_softmax = lambda logits, tau=1.0, hard=False, dim=-1: gumbel_softmax_with_noise(logits, gumbel, tau, hard)
    try:
        ret = model(pts, temperature=temperature, hard=False)
        loss_recon, loss_klv = model.get_loss(ret, pts)
        loss = loss_recon + DVAE_KLD_WEIGHT * loss_klv
        loss.backward()
    finally:
        dvae.F.gumbel_softmax = orig
    return model, ret, loss_recon, loss_klv, loss


def gen_dvae_step():
    """BASELINE config 3 (Stage-I autoencoder step) at B=2: outputs, both losses and parameter gradients."""
    pts = synthetic_clouds(2, 1024, seed=23)
    model, ret, l1, l2, loss = run_reference_dvae(pts)
    whole_coarse, whole_fine, coarse, fine, nb, logits = ret
    res = {"pts": pts.numpy(), "coarse": coarse.detach().numpy(), "fine": fine.detach().numpy(),
           "whole_fine": whole_fine.numpy(), "logits_sample": logits.detach()[:, ::8, ::64].numpy(),
           "logits_absmean": np.float32(logits.abs().mean().item()),
           "loss_recon": np.float32(l1.item()), "loss_klv": np.float32(l2.item()), "loss": np.float32(loss.item())}
    keep_full = ("decoder.mlp.4.bias", "decoder.final_conv.6.weight", "decoder.final_conv.0.weight",
                 "dgcnn_2.layer5.1.weight", "dgcnn_2.input_trans.bias", "dgcnn_1.layer1.1.bias",
                 "encoder.second_conv.3.bias", "encoder.first_conv.0.weight")
    names, norms = [], []
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        names.append(k)
        norms.append(p.grad.norm().item())
        if k in keep_full:
            res["grad/" + k] = p.grad.numpy()
    res["grad/codebook_rows"] = model.codebook.grad[::512].numpy()
    res["grad_names"] = np.array(names)
    res["grad_norms"] = np.array(norms, np.float64)
    for k, b in model.named_buffers():
        if "running" in k:
            res["buf/" + k] = b.numpy()
    np.savez_compressed(os.path.join(GOLD, "dvae_step.npz"), **res)
    print("dvae_step.npz recon", l1.item(), "klv", l2.item(), "n grads", len(names))


def dvae_smooth_weights(coarse_shape, fine_shape, seed=43):
    """Fixed weights of the smooth surrogate loss sum(coarse * Rc) + sum(fine * Rf) (shared with the GPU test)."""
    rng = np.random.default_rng(seed)
    rc = torch.from_numpy(rng.standard_normal(coarse_shape).astype(np.float32)) / float(np.prod(coarse_shape[:-1]))
    rf = torch.from_numpy(rng.standard_normal(fine_shape).astype(np.float32)) / float(np.prod(fine_shape[:-1]))
    return rc, rf


def gen_dvae_step_smooth():
    """The same Stage-I forward as gen_dvae_step, differentiated through a SMOOTH loss (a fixed random linear functional of
    the coarse and fine reconstructions + kld_weight * KL) instead of Chamfer-L1.  Chamfer-L1's gradient is a sum of unit
    vectors towards arg-min partners: piecewise constant in the forward values, so two fp32-grade forwards that differ in
    the last bits re-assign a few partners and the parameter gradients move by a few per cent whatever the backward
    arithmetic does.  This fixture isolates the backward arithmetic of the whole step from that discontinuity."""
    import models.dvae as dvae
    from .ref_dvae import gumbel_softmax_with_noise
    pts = synthetic_clouds(2, 1024, seed=23)
    cfg = shims.easydict(dict(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256,
                              tokens_dims=256, decoder_dims=256))
    model = dvae.DiscreteVAE(cfg)
    fill_params(model, seed=8)
    model.train()
    gumbel = dvae_noise(2, 64, 8192)
    orig = dvae.F.gumbel_softmax
    dvae.F.gumbel_softmax = lambda logits, tau=1.0, hard=False, dim=-1: gumbel_softmax_with_noise(logits, gumbel, tau, hard)
    try:
        ret = model(pts, temperature=1.0, hard=False)
        whole_coarse, whole_fine, coarse, fine, nb, logits = ret
        _, loss_klv = model.get_loss(ret, pts)
        rc, rf = dvae_smooth_weights(tuple(coarse.shape), tuple(fine.shape))
        loss = (coarse * rc).sum() + (fine * rf).sum() + DVAE_KLD_WEIGHT * loss_klv
        loss.backward()
    finally:
        dvae.F.gumbel_softmax = orig
    keep_full = ("decoder.mlp.4.bias", "decoder.final_conv.6.weight", "decoder.final_conv.0.weight",
                 "dgcnn_2.layer5.1.weight", "dgcnn_2.input_trans.bias", "dgcnn_1.layer1.1.bias",
                 "encoder.second_conv.3.bias", "encoder.first_conv.0.weight")
    res = {"loss": np.float32(loss.item())}
    names, norms = [], []
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        names.append(k)
        norms.append(p.grad.norm().item())
        if k in keep_full:
            res["grad/" + k] = p.grad.numpy()
    res["grad/codebook_rows"] = model.codebook.grad[::512].numpy()
    res["grad_names"] = np.array(names)
    res["grad_norms"] = np.array(norms, np.float64)
    np.savez_compressed(os.path.join(GOLD, "dvae_step_smooth.npz"), **res)
    print("dvae_step_smooth.npz loss", loss.item(), "n grads", len(names))


def gen_point_transformer():
    """the unmodified reference PointTransformer (act.py:727-910), transfer_type 'full' (mlp-3 head) and
    'side': eval logits, and one train-mode forward/backward with cross-entropy (head dropout p set to 0 so that the run
    is deterministic; everything else as constructed)."""
    import models.act as act
    res = {}
    pts = synthetic_clouds(4, 1024, seed=29)
    gt = torch.tensor([3, 17, 0, 39])
    res["pts"], res["gt"] = pts.numpy(), gt.numpy()
    for tt in ("full", "side", "linear"):
        cfg = shims.easydict(dict(NAME="PointTransformer", embed_dim=384, depth=12, drop_path_rate=0.0, cls_dim=40,
                                  num_heads=6, group_size=32, num_group=64, encoder_dims=384, transfer_type=tt))
        model = fill_params(act.PointTransformer(cfg), seed=9)
        if tt == "linear":       # the linear head has no 4-sample BatchNorm in front of the loss: a well-conditioned
            for p_ in model.parameters():      # gradient check of the whole backbone -> un-freeze it for this run
                p_.requires_grad = True
        for m in model.cls_head_finetune:
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        model.eval()
        with torch.no_grad():
            res[tt + "/logits_eval"] = model(pts).numpy()
        model.train()
        ret = model(pts)
        loss, acc = model.get_loss_acc(ret, gt)
        loss.backward()
        res[tt + "/logits_train"] = ret.detach().numpy()
        res[tt + "/loss"] = np.float32(loss.item())
        names, norms = [], []
        for k, p in model.named_parameters():
            if p.grad is not None:
                names.append(k)
                norms.append(p.grad.norm().item())
        res[tt + "/grad_names"] = np.array(names)
        res[tt + "/grad_norms"] = np.array(norms, np.float64)
        last = len(model.cls_head_finetune) - 1
        res[tt + f"/grad/cls_head_finetune.{last}.weight"] = model.cls_head_finetune[last].weight.grad.numpy()
        res[tt + "/grad/cls_token"] = model.cls_token.grad.numpy()
        if tt != "side":          # 'side' freezes everything without 'side' / 'cls' in its name (act.py:797-806)
            res[tt + "/grad/blocks.blocks.0.attn.qkv.weight"] = model.blocks.blocks[0].attn.qkv.weight.grad[::16, ::8].numpy()
        else:
            res[tt + "/grad/side_projection.weight"] = model.side_projection.weight.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "point_transformer.npz"), **res)
    print("point_transformer.npz", {k: float(res[k + "/loss"]) for k in ("full", "side", "linear")})


def main():
    if not os.path.isdir(shims.REFERENCE_ROOT):
        sys.exit("needs /root/reference (authoring container only)")
    shims.install()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    gens = {"group": gen_group, "block_cfg1": gen_block_cfg1, "encoder": gen_encoder, "student_step": gen_student_step,
            "teacher": gen_teacher, "dvae_step": gen_dvae_step, "dvae_step_smooth": gen_dvae_step_smooth,
            "point_transformer": gen_point_transformer}
    for name in (sys.argv[1:] or list(gens)):      # `python -m oracle.make_golden [fixture ...]`: all by default
        gens[name]()


if __name__ == "__main__":
    main()
