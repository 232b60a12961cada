"""Registry-compatible model classes of the hot path (same NAME, constructor contract `cls(cfg)` and
state_dict keys as the reference: /root/reference/models/act.py:1099-1258, utils/registry.py:272-285).

`ACT_PointDistillation` is the Stage-II model that tools/runner_pretrain.py:139 calls as
`loss = base_model(points)`.  Like the reference (act.py:1151-1160) `cls(cfg)` builds the frozen Stage-I teacher
`dvae_tokenizer` (ACTPromptedDiscreteVAEwithVIT on the act_b200 kernels, act_b200/teacher.py), strict-loads
`cfg.dvae_config.ckpt` into it and freezes it.  Only on explicit request is something else used: `teacher="synthetic"`
(a deterministic parameter-free target: student-only timing / parity runs) or any callable
(neighborhood, center) -> [B,G,C] features.
"""
import torch
import torch.nn as nn

from . import layers, ops
from .modules import Encoder, Group, TransformerDecoder, TransformerEncoder, VisableOnlyMaskTransformer, pos_mlp

MODELS = {}


def register(cls):
    MODELS[cls.__name__] = cls
    return cls


class Cfg(dict):
    """Minimal EasyDict look-alike (the reference passes easydict.EasyDict built from cfgs/*.yaml)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = Cfg(v) if isinstance(v, dict) else v

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def default_config(mask_ratio=0.6, drop_path_rate=0.1, num_group=64, group_size=32, depth=12, decoder_depth=2):
    """model: block of cfgs/pretrain/pretrain_act_distill.yaml (BASELINE config 2 uses mask_ratio 0.6)."""
    return Cfg(NAME="ACT_PointDistillation", loss="cosine",
               transformer_config=dict(mask_ratio=mask_ratio, mask_type="rand", proj="linear", embed_dim=384,
                                       encoder_dims=384, depth=depth, drop_path_rate=drop_path_rate, cls_dim=512,
                                       replace_pob=0.0, num_heads=6, decoder_depth=decoder_depth,
                                       decoder_num_heads=6, return_all_tokens=False, cls_loss=False,
                                       register_shallow_hook=9),
               dvae_config=dict(num_group=num_group, group_size=group_size, encoder_dims=384, num_tokens=8192,
                                tokens_dims=384, decoder_dims=384, ckpt=None))


class SyntheticTeacher(nn.Module):
    """Stand-in for dvae_tokenizer.forward_tokenizer_features: fixed pseudo-random features that depend on the
    group centres (so different clouds get different targets), no parameters, no gradient."""

    def __init__(self, dim):
        super().__init__()
        g = torch.Generator().manual_seed(1234)
        self.register_buffer("proj", torch.randn(3, dim, generator=g), persistent=False)
        self.register_buffer("phase", torch.rand(dim, generator=g) * 6.2831853, persistent=False)

    @torch.no_grad()
    def forward(self, neighborhood, center):
        return torch.sin(center @ self.proj * 3.0 + self.phase)


@register
class ACT_PointDistillation(nn.Module):
    def __init__(self, config, teacher=None):
        super().__init__()
        self.config = config
        tc, dc = config.transformer_config, config.dvae_config
        if tc.cls_loss or tc.proj != "linear" or config.loss not in ("cosine", "l2", "smoothl1"):
            raise NotImplementedError("act_b200 implements cls_loss False, proj linear, loss cosine (the shipped config) / "
                                      "l2 / smoothl1")
        self.loss_type = config.loss
        self.mask_ratio = tc.mask_ratio
        self.embed_dim = tc.embed_dim
        self.ACT_encoder = VisableOnlyMaskTransformer(config)
        self.group_size, self.num_group = dc.group_size, dc.num_group
        self.drop_path_rate = tc.drop_path_rate
        self.decoder_depth, self.decoder_num_heads = tc.decoder_depth, tc.decoder_num_heads
        # The tokenizer contract of the reference (act.py:1151-1160, build_tokenizer): construct the prompted dVAE + ViT
        # teacher under the attribute name `dvae_tokenizer` (so a Stage-II state_dict has the reference's
        # `dvae_tokenizer.*` keys), strict-load `dvae_config.ckpt`, freeze every parameter.  `teacher=None` / "native"
        # is that default.  The reference cannot be constructed without the checkpoint file; here a missing `ckpt`
        # (None) leaves the teacher at its initialisation -- the offline situation of the benchmark and the tests --
        # while a path that is set but unreadable raises, as torch.load does in the reference.
        if teacher is None or teacher == "native":
            from .teacher import ACTPromptedDiscreteVAEwithVIT
            self.dvae_tokenizer = ACTPromptedDiscreteVAEwithVIT(dc)
            ckpt_path = dc.get("ckpt") if isinstance(dc, dict) else getattr(dc, "ckpt", None)
            if ckpt_path:
                ckpt = torch.load(ckpt_path, map_location="cpu")
                base_ckpt = {k.replace("module.", ""): v for k, v in ckpt["base_model"].items()}
                self.dvae_tokenizer.load_state_dict(base_ckpt, strict=True)
            for p in self.dvae_tokenizer.parameters():
                p.requires_grad = False
            teacher = self.dvae_tokenizer.forward_tokenizer_features
        elif teacher == "synthetic":
            teacher = SyntheticTeacher(dc.tokens_dims)
        elif not callable(teacher):
            raise ValueError("teacher: None / 'native' (the reference's frozen dvae_tokenizer), 'synthetic', or a callable")
        object.__setattr__(self, "teacher", teacher)      # not a registered submodule (dvae_tokenizer above already is)
        self.group_divider = Group(num_group=self.num_group, group_size=self.group_size)
        self.proj_head = nn.Linear(self.embed_dim, dc.tokens_dims)
        if self.mask_ratio > 0.:
            self.mask_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
            self.decoder_pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, self.embed_dim))
            dpr = [x.item() for x in torch.linspace(0, self.drop_path_rate, self.decoder_depth)]
            self.ACT_decoder = TransformerDecoder(embed_dim=self.embed_dim, depth=self.decoder_depth,
                                                  drop_path_rate=dpr, num_heads=self.decoder_num_heads)
            nn.init.trunc_normal_(self.mask_token, std=.02)
        else:
            raise NotImplementedError("mask_ratio 0 (no decoder) is not part of the hot path")

    # trainable parameters that never get a gradient in the distillation step (act.py:187-196: token-classification
    # and contrast heads of the Point-BERT lineage); pass to layers.FlatParams(exclude=...)
    UNUSED_PARAMETERS = ("ACT_encoder.lm_head.", "ACT_encoder.cls_head.")

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        if isinstance(self.teacher, nn.Module):
            self.teacher._apply(fn)
        return self

    def forward_eval(self, pts):
        with torch.no_grad():
            neighborhood, center = self.group_divider(pts)
            return self.ACT_encoder(neighborhood, center, only_cls_tokens=True, noaug=True)

    def forward(self, pts, noaug=False, mask=None, teacher_feat=None, group=None, **kwargs):
        """group (optional): a precomputed (neighborhood, center) of `pts` -- engine.PretrainStep's pipelined mode runs the
        tokenizer and the frozen teacher ahead of the student (they do not depend on the student's weights)."""
        if noaug:
            return self.forward_eval(pts)
        neighborhood, center = self.group_divider(pts) if group is None else group
        # The frozen teacher's forward (act.py:1216-1217) depends only on the tokenizer's output, and the student's
        # encoder/decoder forward is a chain of small latency-bound kernels that leave most SMs idle: fork the teacher
        # onto a second stream right after the Group tokenizer and join before the loss, so inside the captured graph
        # the two branches run concurrently.
        fork = None
        if teacher_feat is None:
            fork = layers._SideStream(pts.device, tag="teacher")
            box = []

            def run_teacher():
                with torch.no_grad():
                    box.append(self.teacher(neighborhood, center))
            fork.run(run_teacher, neighborhood, center)
        student, order, n_vis = self.forward_student(neighborhood, center, mask)
        if fork is not None:
            fork.join()
            teacher_feat = box[0]
            if fork.side is not None:
                teacher_feat.record_stream(fork.main)
        return self.distill_loss(student, teacher_feat, order, n_vis)

    def forward_student(self, neighborhood, center, mask=None):
        """The trainable half of forward() up to the projection head (act.py:1212-1228): -> (student [B,num_mask,C],
        order [B,G] = visible groups first, n_vis).  Does not need the teacher's features."""
        x_vis, mask, ex = self.ACT_encoder(neighborhood, center, mask=mask, return_extras=True)
        B, n_vis, C = x_vis.shape
        G = center.shape[1]
        num_mask = G - n_vis
        order = ex["order"]                                    # visible groups first, then masked, original order
        pos_full = pos_mlp(self.decoder_pos_embed, ex["centers_sorted"])          # [pos(vis) | pos(mask)]  [B,G,C]
        # cat([x_vis, mask_token.expand]) read straight from the encoder output past its cls row (no slice copy)
        x_full = layers.assemble_rows(ex["encoded"], self.mask_token, B, n_vis, G, False, src_off=1)
        x_dec = self.ACT_decoder(x_full, pos_full, num_mask)
        student = layers.linear(x_dec, self.proj_head.weight, self.proj_head.bias)
        return student, order, n_vis

    def distill_loss(self, student, teacher_feat, order, n_vis):
        """act.py:1229-1254: the teacher's features at the masked groups against the student's predictions."""
        G = order.shape[1]
        teacher = ops.gather_rows(teacher_feat.detach(), order, n_vis, G - n_vis)  # teacher_feat[mask], original order
        if self.loss_type == "cosine":
            return layers.cosine_loss(student, teacher)
        return layers.pointwise_loss(student, teacher.reshape(student.shape), self.loss_type)    # act.py:1255-1256


@register
class PointTransformer(nn.Module):
    """models/act.py:727-910: the fine-tune / inference classifier on the hot-path kernels -- Group ->
    mini-PointNet -> cls token + all G tokens through the Blocks -> LayerNorm -> cat(cls, max over tokens) -> head.  Same
    constructor contract (`cls(config)` with embed_dim, depth, drop_path_rate, cls_dim, num_heads, group_size, num_group,
    encoder_dims, transfer_type), attribute names (=> state_dict keys) and transfer-type freezing rules; trainable
    (every piece has its backward).  The tiny [B, 2C] classification head stays on PyTorch ops."""

    def __init__(self, config, **kwargs):
        super().__init__()
        self.config = config
        self.embed_dim, self.depth = config.embed_dim, config.depth
        self.drop_path_rate, self.cls_dim, self.num_heads = config.drop_path_rate, config.cls_dim, config.num_heads
        self.group_size, self.num_group, self.encoder_dims = config.group_size, config.num_group, config.encoder_dims
        self.group_divider = Group(num_group=self.num_group, group_size=self.group_size)
        self.encoder = Encoder(encoder_channel=self.encoder_dims)
        self.reduce_dim = (nn.Linear(self.encoder_dims, self.embed_dim) if self.encoder_dims != self.embed_dim
                           else nn.Identity())
        self.cls_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
        self.cls_pos = nn.Parameter(torch.randn(1, 1, self.embed_dim))
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, self.embed_dim))
        dpr = [x.item() for x in torch.linspace(0, self.drop_path_rate, self.depth)]
        self.blocks = TransformerEncoder(embed_dim=self.embed_dim, depth=self.depth, drop_path_rate=dpr,
                                         num_heads=self.num_heads)
        self.norm = nn.LayerNorm(self.embed_dim)
        tt = config.transfer_type
        if tt == 'linear':
            self.cls_head_finetune = nn.Sequential(nn.Linear(self.embed_dim * 2, self.cls_dim))
        else:
            self.cls_head_finetune = nn.Sequential(
                nn.Linear(self.embed_dim * 2, 256), nn.BatchNorm1d(256), nn.ReLU(inplace=True), nn.Dropout(0.5),
                nn.Linear(256, 256), nn.BatchNorm1d(256), nn.ReLU(inplace=True), nn.Dropout(0.5),
                nn.Linear(256, self.cls_dim))
        self.loss_ce = nn.CrossEntropyLoss()
        self.side = None
        if tt == "side":                                            # act.py:808-814
            self.side_alpha = nn.Parameter(torch.Tensor([0.0]))
            self.side = Encoder(encoder_channel=self.embed_dim)
            self.side_projection = nn.Linear(self.embed_dim, self.embed_dim, bias=False)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        nn.init.trunc_normal_(self.cls_pos, std=.02)
        if tt != 'full':                                            # act.py:795-806
            for name, param in self.named_parameters():
                if tt in ('mlp-3', 'linear'):
                    keep = 'cls' in name
                elif tt == 'side':
                    keep = 'side' in name or 'cls' in name
                elif tt == 'bit-fit':
                    keep = 'bias' in name or 'cls' in name
                else:
                    keep = True
                if not keep:
                    param.requires_grad = False

    def load_model_from_ckpt(self, bert_ckpt_path, custom_loading=False):
        """models/act.py:829-867 (what tools/runner_finetune.py calls): load a Stage-II checkpoint into the classifier --
        strip `module.`, map `ACT_encoder.*` / `base_model.*` onto this module's keys, load non-strictly; returns the
        incompatible-keys record the reference logs.  `None` re-initialises like the reference ("training from scratch")."""
        if bert_ckpt_path is None:
            self.apply(self._init_weights)
            return None
        ckpt = torch.load(bert_ckpt_path, map_location="cpu")
        if custom_loading:
            base_ckpt = {k.replace("module.point_encoder.", ""): v for k, v in ckpt['state_dict'].items()}
            base_ckpt = {k.replace("encoder", "blocks"): v for k, v in base_ckpt.items()}
            base_ckpt = {k.replace("patch_embed", "encoder"): v for k, v in base_ckpt.items()}
        else:
            base_ckpt = {k.replace("module.", ""): v for k, v in ckpt['base_model'].items()}
        for k in list(base_ckpt.keys()):
            if k.startswith('ACT_encoder'):
                base_ckpt[k[len('ACT_encoder.'):]] = base_ckpt[k]
                del base_ckpt[k]
            elif k.startswith('base_model'):
                base_ckpt[k[len('base_model.'):]] = base_ckpt[k]
                del base_ckpt[k]
        incompatible = self.load_state_dict(base_ckpt, strict=False)
        self.last_incompatible_keys = incompatible
        return incompatible

    def _init_weights(self, m):
        if isinstance(m, (nn.Linear, nn.Conv1d)):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def get_loss_acc(self, ret, gt):
        loss = self.loss_ce(ret, gt.long())
        pred = ret.argmax(-1)
        acc = (pred == gt).sum() / float(gt.size(0))
        return loss, acc * 100

    def forward(self, pts):
        neighborhood, center = self.group_divider(pts)
        tokens = self.encoder(neighborhood)                                    # B G C
        if not isinstance(self.reduce_dim, nn.Identity):
            tokens = layers.linear(tokens, self.reduce_dim.weight, self.reduce_dim.bias)
        B, G = tokens.shape[:2]
        x = layers.assemble_rows(tokens, self.cls_token, B, G, G + 1, True)               # cat(cls_token, tokens)
        pos = layers.assemble_rows(pos_mlp(self.pos_embed, center), self.cls_pos, B, G, G + 1, True)
        x = self.blocks(x, pos)
        x = layers.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        if self.side is not None:
            side = layers.linear(self.side(neighborhood), self.side_projection.weight)
            a = torch.sigmoid(self.side_alpha)
            side = a * x[:, 1:] + (1 - a) * side
            concat_f = torch.cat([x[:, 0], side.max(1)[0]], dim=-1)
        else:
            concat_f = torch.cat([x[:, 0], x[:, 1:].max(1)[0]], dim=-1)
        return self.cls_head_finetune(concat_f)


def build_model_from_cfg(cfg, **kwargs):
    """models/build.py:7-17 equivalent: look up cfg.NAME, call cls(cfg)."""
    return MODELS[cfg.NAME](cfg, **kwargs)
