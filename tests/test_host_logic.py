"""CPU: host-side logic of the boundary that needs no kernel -- the optimizer state's checkpoint layout, the reference's
tokenizer contract in ACT_PointDistillation.__init__, PointTransformer.load_model_from_ckpt."""
import os

import pytest
import torch
import torch.nn as nn

from act_b200 import layers, models


def _no_decay(n, p):
    return p.dim() <= 1 or n.endswith(".bias") or "token" in n


def _reference_adamw(module, lr=1e-3, wd=0.05):
    """tools/builder.py:37-55 (add_weight_decay + optim.AdamW): group 0 = no-decay, group 1 = decay."""
    named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
    return torch.optim.AdamW([{"params": [p for n, p in named if _no_decay(n, p)], "weight_decay": 0.0},
                              {"params": [p for n, p in named if not _no_decay(n, p)], "weight_decay": wd}], lr=lr)


def _net():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(8, 16), nn.LayerNorm(16), nn.Linear(16, 8))


def test_flatparams_state_dict_is_torch_adamw_layout_and_round_trips():
    m = _net()
    fp = layers.FlatParams(m, lr=2e-3, weight_decay=0.05)
    fp.step_count = 7
    fp.exp_avg.normal_()
    fp.exp_avg_sq.uniform_()
    sd = fp.state_dict()
    opt = _reference_adamw(_net())
    opt.load_state_dict(sd)                                    # the reference's resume_optimizer (builder.py:119-130)
    plist = [p for g in opt.param_groups for p in g["params"]]
    for i, p in enumerate(plist):
        assert torch.equal(opt.state[p]["exp_avg"], sd["state"][i]["exp_avg"])
        assert float(opt.state[p]["step"]) == 7.0 and opt.state[p]["exp_avg"].shape == p.shape
    assert opt.param_groups[0]["weight_decay"] == 0.0 and opt.param_groups[1]["weight_decay"] == 0.05
    assert opt.param_groups[0]["lr"] == 2e-3
    fp2 = layers.FlatParams(_net())
    fp2.load_state_dict(opt.state_dict())                      # and back, from a genuine torch.optim.AdamW state_dict
    assert torch.equal(fp2.exp_avg, fp.exp_avg) and torch.equal(fp2.exp_avg_sq, fp.exp_avg_sq)
    assert fp2.step_count == 7 and fp2.lr == 2e-3 and fp2.weight_decay == 0.05


def test_flatparams_excluded_parameters_keep_their_index_but_get_no_state():
    m = nn.ModuleDict({"a": nn.Linear(8, 8), "unused": nn.Linear(8, 8), "b": nn.Linear(8, 8)})
    fp = layers.FlatParams(m, exclude=("unused.",))
    fp.step_count = 1
    sd = fp.state_dict()
    n_all = sum(1 for _ in m.parameters())
    assert sorted(i for g in sd["param_groups"] for i in g["params"]) == list(range(n_all))
    assert len(sd["state"]) == n_all - 2                       # torch creates no state for a parameter without .grad
    opt = _reference_adamw(m)
    opt.load_state_dict(sd)


def test_flatparams_lr_scheduler_hook():
    fp = layers.FlatParams(_net(), lr=1e-3)
    for g in fp.param_groups:                                  # what timm's / torch's schedulers do every epoch
        g["lr"] = 5e-4
    assert fp.lr == 5e-4
    fp.set_hyper()
    assert abs(fp.hyper[0].item() - 5e-4) < 1e-10 and fp.step_count == 1
    fp.param_groups[0]["lr"] = 1e-4
    with pytest.raises(ValueError):
        fp.set_hyper()


def test_distillation_model_builds_loads_and_freezes_the_teacher_like_the_reference(tmp_path):
    """models/act.py:1151-1160: ACTPromptedDiscreteVAEwithVIT under `dvae_tokenizer`, strict load of dvae_config.ckpt,
    requires_grad False everywhere in it; a checkpoint with a missing key must raise (strict=True)."""
    cfg = models.default_config()
    donor = models.ACT_PointDistillation(cfg)                  # ckpt None: teacher stays at its initialisation
    assert hasattr(donor, "dvae_tokenizer") and all(not p.requires_grad for p in donor.dvae_tokenizer.parameters())
    assert any(k.startswith("dvae_tokenizer.visual_embed.0.11.") for k in donor.state_dict())
    sd = {("module." + k): v + 1.0 if v.dtype.is_floating_point else v for k, v in donor.dvae_tokenizer.state_dict().items()}
    path = os.path.join(tmp_path, "dvae.pth")
    torch.save({"base_model": sd}, path)
    cfg.dvae_config.ckpt = path
    model = models.ACT_PointDistillation(cfg)
    for k, v in model.dvae_tokenizer.state_dict().items():
        assert torch.equal(v, sd["module." + k]), k
    assert all(not p.requires_grad for p in model.dvae_tokenizer.parameters())
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert trainable and not any(n.startswith("dvae_tokenizer") for n in trainable)
    sd.pop("module.codebook")
    torch.save({"base_model": sd}, path)
    with pytest.raises(RuntimeError):
        models.ACT_PointDistillation(cfg)
    cfg.dvae_config.ckpt = os.path.join(tmp_path, "absent.pth")
    with pytest.raises(FileNotFoundError):
        models.ACT_PointDistillation(cfg)
    assert not hasattr(models.ACT_PointDistillation(models.default_config(), teacher="synthetic"), "dvae_tokenizer")


def test_point_transformer_load_model_from_ckpt(tmp_path):
    """models/act.py:829-867: a Stage-II checkpoint ('base_model', optional 'module.' prefix, 'ACT_encoder.' keys) lands
    on the classifier's encoder / blocks / norm / tokens; the fine-tune head is the only thing missing."""
    student = models.ACT_PointDistillation(models.default_config(), teacher="synthetic")
    path = os.path.join(tmp_path, "ckpt.pth")
    torch.save({"base_model": {"module." + k: v for k, v in student.state_dict().items()}}, path)
    cfg = models.Cfg(NAME="PointTransformer", embed_dim=384, depth=12, drop_path_rate=0.1, cls_dim=40, num_heads=6,
                     group_size=32, num_group=64, encoder_dims=384, transfer_type="full")
    clf = models.PointTransformer(cfg)
    inc = clf.load_model_from_ckpt(path)
    assert all(k.startswith("cls_head_finetune") for k in inc.missing_keys), inc.missing_keys
    for k in ("encoder.first_conv.0.weight", "blocks.blocks.11.mlp.fc2.weight", "norm.weight", "cls_token", "pos_embed.2.bias"):
        assert torch.equal(clf.state_dict()[k], student.state_dict()["ACT_encoder." + k]), k
    assert any(k.startswith("ACT_decoder") for k in inc.unexpected_keys)
    assert clf.load_model_from_ckpt(None) is None
