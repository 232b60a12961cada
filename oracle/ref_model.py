"""Plain-PyTorch fp32 CPU restatement of the reference's Stage-II student path.
TEST INFRASTRUCTURE ONLY (oracle): imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by act_b200/.

Why it exists: the reference is Python and cannot travel to the GPU box (/root/reference is absent
there), and it needs nine uninstalled packages.  This file restates, module by module, exactly what
the reference computes, with the same class layout and state_dict keys, so that
  * tests here (CPU container) can check it against the REAL reference imported through
    oracle/shims.py (tests/test_oracle_vs_reference.py), and
  * on the GPU box it is the checker for the CUDA path and the timed CPU baseline.
PARITY PINNING: pinned against the unmodified reference modules (models/dvae.py Encoder/Group,
models/act.py Block/TransformerEncoder/VisableOnlyMaskTransformer/ACT_PointDistillation) run in
this container, and against the golden fixtures under tests/golden/ generated from them by
oracle/make_golden.py.  The native ops underneath Group (FPS / kNN) are the C restatement
oracle/cpu_ref.c (parity unpinned against the upstream CUDA binaries, see its header).

Each class cites the reference lines it follows.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cpu_ref


class Group(nn.Module):
    """models/dvae.py:154-183 (Group) + utils/misc.py:39-46 (fps)."""

    def __init__(self, num_group, group_size):
        super().__init__()
        self.num_group, self.group_size = num_group, group_size

    @torch.no_grad()
    def forward(self, xyz):
        nb, center, idx, fps_idx = cpu_ref.group(xyz.detach().cpu().numpy(), self.num_group, self.group_size)
        self.last_idx, self.last_fps_idx = idx, fps_idx
        return torch.from_numpy(nb), torch.from_numpy(center)


class Encoder(nn.Module):
    """models/dvae.py:185-215 (mini-PointNet; BatchNorm1d in whatever mode the module is in)."""

    def __init__(self, encoder_channel):
        super().__init__()
        self.encoder_channel = encoder_channel
        self.first_conv = nn.Sequential(nn.Conv1d(3, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
                                        nn.Conv1d(128, 256, 1))
        self.second_conv = nn.Sequential(nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
                                         nn.Conv1d(512, encoder_channel, 1))

    def forward(self, point_groups):
        bs, g, n, _ = point_groups.shape
        pg = point_groups.reshape(bs * g, n, 3)
        f = self.first_conv(pg.transpose(2, 1))
        fg = torch.max(f, dim=2, keepdim=True)[0]
        f = torch.cat([fg.expand(-1, -1, n), f], dim=1)
        f = self.second_conv(f)
        fg = torch.max(f, dim=2, keepdim=False)[0]
        return fg.reshape(bs, g, self.encoder_channel)


class Mlp(nn.Module):
    """models/act.py:25-42."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Attention(nn.Module):
    """models/act.py:45-69 (qkv without bias, scale = head_dim**-0.5, materialised softmax)."""

    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class Block(nn.Module):
    """models/act.py:72-90 with drop_path = 0 (parity runs use drop_path_rate 0; DropPath is an
    RNG-dependent per-sample gate, timm 0.5.4)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.attn = Attention(dim, num_heads)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class TransformerEncoder(nn.Module):
    """models/act.py:93-112: x = block(x + pos) at EVERY layer."""

    def __init__(self, embed_dim, depth, num_heads):
        super().__init__()
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads) for _ in range(depth)])

    def forward(self, x, pos):
        for blk in self.blocks:
            x = blk(x + pos)
        return x


class TransformerDecoder(nn.Module):
    """models/act.py:115-145: blocks, then LayerNorm on the last `return_token_num` tokens."""

    def __init__(self, embed_dim, depth, num_heads):
        super().__init__()
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim)

    def forward(self, x, pos, return_token_num):
        for blk in self.blocks:
            x = blk(x + pos)
        return self.norm(x[:, -return_token_num:])


def mask_center_rand(B, G, mask_ratio):
    """models/act.py:244-267 (_mask_center_rand): host numpy RNG, exactly int(ratio*G) ones per cloud."""
    num_mask = int(mask_ratio * G)
    overall = np.zeros([B, G])
    for i in range(B):
        m = np.hstack([np.zeros(G - num_mask), np.ones(num_mask)])
        np.random.shuffle(m)
        overall[i, :] = m
    return torch.from_numpy(overall).to(torch.bool)


class VisableOnlyMaskTransformer(nn.Module):
    """models/act.py:148-309 (student encoder; reduce_dim is Identity when encoder_dims == embed_dim)."""

    def __init__(self, embed_dim=384, depth=12, num_heads=6, encoder_dims=384, mask_ratio=0.6, num_tokens=8192,
                 cls_dim=512):
        super().__init__()
        self.mask_ratio = mask_ratio
        self.encoder = Encoder(encoder_dims)
        self.reduce_dim = nn.Linear(encoder_dims, embed_dim) if encoder_dims != embed_dim else nn.Identity()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.cls_pos = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, embed_dim))
        self.blocks = TransformerEncoder(embed_dim, depth, num_heads)
        self.norm = nn.LayerNorm(embed_dim)
        self.lm_head = nn.Linear(embed_dim, num_tokens)          # unused by the distillation loss
        self.cls_head = nn.Sequential(nn.Linear(embed_dim, cls_dim), nn.GELU(), nn.Linear(cls_dim, cls_dim))

    def forward(self, neighborhood, center, mask=None):
        B, G, _ = center.shape
        if mask is None:
            mask = mask_center_rand(B, G, self.mask_ratio)
        tokens = self.reduce_dim(self.encoder(neighborhood))
        C = tokens.shape[-1]
        x_vis = tokens[~mask].reshape(B, -1, C)
        pos = self.pos_embed(center[~mask].reshape(B, -1, 3))
        x_vis = torch.cat((self.cls_token.expand(B, -1, -1), x_vis), dim=1)
        pos = torch.cat((self.cls_pos.expand(B, -1, -1), pos), dim=1)
        x_vis = self.norm(self.blocks(x_vis, pos))
        return x_vis[:, 1:], mask


class ACTPointDistillationStudent(nn.Module):
    """models/act.py:1099-1258 (ACT_PointDistillation.forward) with the frozen teacher's output
    `teacher_feat[B,G,C]` supplied by the caller (act.py:1216-1217 is SURVEY row f1, "next").
    cls_loss False, proj 'linear', loss 'cosine' (cfgs/pretrain/pretrain_act_distill.yaml)."""

    def __init__(self, num_group=64, group_size=32, embed_dim=384, depth=12, num_heads=6, decoder_depth=2,
                 decoder_num_heads=6, mask_ratio=0.6, tokens_dims=384):
        super().__init__()
        self.ACT_encoder = VisableOnlyMaskTransformer(embed_dim, depth, num_heads, embed_dim, mask_ratio)
        self.group_divider = Group(num_group, group_size)
        self.proj_head = nn.Linear(embed_dim, tokens_dims)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.decoder_pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, embed_dim))
        self.ACT_decoder = TransformerDecoder(embed_dim, decoder_depth, decoder_num_heads)

    def forward(self, pts, teacher_feat, mask=None):
        neighborhood, center = self.group_divider(pts)
        x_vis, mask = self.ACT_encoder(neighborhood, center, mask)
        B, _, C = x_vis.shape
        pos_vis = self.decoder_pos_embed(center[~mask]).reshape(B, -1, C)
        pos_mask = self.decoder_pos_embed(center[mask]).reshape(B, -1, C)
        num_mask = pos_mask.shape[1]
        x_full = torch.cat([x_vis, self.mask_token.expand(B, num_mask, -1)], dim=1)
        pos_full = torch.cat([pos_vis, pos_mask], dim=1)
        student = self.proj_head(self.ACT_decoder(x_full, pos_full, num_mask))
        teacher = teacher_feat[mask].reshape(B, -1, student.shape[-1])
        # act.py:1243-1254: sum_b (1 - mean_tok cos(student[b], teacher[b])) / B, cosine eps 1e-8
        loss = torch.zeros(1)
        for b in range(B):
            loss = loss + (1 - F.cosine_similarity(student[b], teacher[b], 1, 1e-8).mean())
        return loss.mean() / B


def fill_params(module, seed=0):
    """Deterministic, RNG-library-independent parameter fill used by BOTH the golden generator (on the
    real reference modules) and the parity tests (on the oracle / CUDA modules): every tensor of the
    state_dict is drawn from numpy's PCG64 seeded by (seed, crc32(key)); so identical keys+shapes give
    identical weights whatever the construction order.  Scales mimic a trained net enough to exercise
    every term (non-zero biases, non-unit norm gains, BN running stats)."""
    import zlib
    sd = module.state_dict()
    out = {}
    for k, v in sd.items():
        rng = np.random.default_rng([seed, zlib.crc32(k.encode())])
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
            continue
        a = rng.standard_normal(tuple(v.shape)).astype(np.float32)
        if k.endswith("running_var"):
            a = 1.0 + 0.1 * np.abs(a)
        elif k.endswith("running_mean"):
            a = 0.05 * a
        elif v.ndim <= 1 and k.endswith("weight"):          # LayerNorm / BatchNorm gains
            a = 1.0 + 0.1 * a
        elif k.endswith("bias"):
            a = 0.02 * a
        elif v.ndim >= 2 and k.endswith("weight"):
            fan_in = int(np.prod(v.shape[1:]))
            a = a * (0.7 / np.sqrt(fan_in))
        else:                                                  # tokens
            a = 0.02 * a
        out[k] = torch.from_numpy(a).reshape(v.shape)
    module.load_state_dict(out, strict=True)
    return module


def synthetic_clouds(B, N, seed=20231017):
    """SURVEY.md 8(d) synthetic ShapeNet-shaped clouds: ellipsoid-surface samples, jitter, unit-sphere
    normalisation (datasets/ShapeNet55Dataset.py:45-67), then PointcloudScaleAndTranslate
    (datasets/data_transforms.py:20-34).  fp32 [B,N,3]."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(B, N, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    radii = 0.3 + 0.7 * torch.rand(B, 1, 3, generator=g)
    p = d * radii + 0.01 * torch.randn(B, N, 3, generator=g)
    p = p - p.mean(dim=1, keepdim=True)
    p = p / p.norm(dim=-1).max(dim=1)[0].view(B, 1, 1)
    scale = 2.0 / 3.0 + (1.5 - 2.0 / 3.0) * torch.rand(B, 1, 3, generator=g)
    trans = -0.2 + 0.4 * torch.rand(B, 1, 3, generator=g)
    return (p * scale + trans).contiguous().float()


def scale_and_translate(pc, scale_low=2. / 3., scale_high=3. / 2., translate_range=0.2):
    """PointcloudScaleAndTranslate.__call__ (datasets/data_transforms.py:20-34) restated for the CPU: per cloud, numpy's
    GLOBAL RNG draws uniform(size=3) scale then uniform(size=3) translation (float64 -> float32), and
    pc[i] = pc[i] * scale + translation in fp32 (multiply, then add).  In place, returns pc."""
    for i in range(pc.shape[0]):
        xyz1 = np.random.uniform(low=scale_low, high=scale_high, size=[3])
        xyz2 = np.random.uniform(low=-translate_range, high=translate_range, size=[3])
        pc[i, :, 0:3] = torch.mul(pc[i, :, 0:3], torch.from_numpy(xyz1).float()) + torch.from_numpy(xyz2).float()
    return pc


class PointTransformer(nn.Module):
    """models/act.py:727-910 (the fine-tune / inference classifier): Group -> Encoder -> cls token + all G
    tokens through the Blocks -> LayerNorm -> cat(cls, max over tokens) -> head.  transfer_type 'linear' uses the linear
    head, every other type the mlp-3 head (act.py:771-789); 'side' adds the side Encoder (act.py:808-814, 899-903)."""

    def __init__(self, embed_dim=384, depth=12, num_heads=6, cls_dim=40, group_size=32, num_group=64, encoder_dims=384,
                 transfer_type="full"):
        super().__init__()
        self.group_divider = Group(num_group, group_size)
        self.encoder = Encoder(encoder_dims)
        self.reduce_dim = nn.Linear(encoder_dims, embed_dim) if encoder_dims != embed_dim else nn.Identity()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.cls_pos = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, embed_dim))
        self.blocks = TransformerEncoder(embed_dim, depth, num_heads)
        self.norm = nn.LayerNorm(embed_dim)
        if transfer_type == "linear":
            self.cls_head_finetune = nn.Sequential(nn.Linear(embed_dim * 2, cls_dim))
        else:
            self.cls_head_finetune = nn.Sequential(
                nn.Linear(embed_dim * 2, 256), nn.BatchNorm1d(256), nn.ReLU(inplace=True), nn.Dropout(0.5),
                nn.Linear(256, 256), nn.BatchNorm1d(256), nn.ReLU(inplace=True), nn.Dropout(0.5), nn.Linear(256, cls_dim))
        self.side = None
        if transfer_type == "side":
            self.side_alpha = nn.Parameter(torch.Tensor([0.0]))
            self.side = Encoder(embed_dim)
            self.side_projection = nn.Linear(embed_dim, embed_dim, bias=False)

    def forward(self, pts):
        neighborhood, center = self.group_divider(pts)
        tokens = self.reduce_dim(self.encoder(neighborhood))
        B = tokens.shape[0]
        x = torch.cat((self.cls_token.expand(B, -1, -1), tokens), dim=1)
        pos = torch.cat((self.cls_pos.expand(B, -1, -1), self.pos_embed(center)), dim=1)
        x = self.norm(self.blocks(x, pos))
        if self.side is not None:
            a = torch.sigmoid(self.side_alpha)
            side = a * x[:, 1:] + (1 - a) * self.side_projection(self.side(neighborhood))
            f = torch.cat([x[:, 0], side.max(1)[0]], dim=-1)
        else:
            f = torch.cat([x[:, 0], x[:, 1:].max(1)[0]], dim=-1)
        return self.cls_head_finetune(f)

    @staticmethod
    def get_loss_acc(ret, gt):
        """act.py:820-824."""
        loss = F.cross_entropy(ret, gt.long())
        acc = (ret.argmax(-1) == gt).sum() / float(gt.size(0))
        return loss, acc * 100


def mask_center_block(center, mask_ratio, index=None):
    """models/act.py:215-243 (_mask_center_block): per cloud, mask the int(ratio * G) centres nearest to a random centre
    (Python's random.randint, consumed one call per cloud like the reference; `index` injects the draws)."""
    import random
    out = []
    for b, points in enumerate(center):
        points = points.unsqueeze(0)
        i = random.randint(0, points.size(1) - 1) if index is None else int(index[b])
        dist = torch.norm(points[:, i].reshape(1, 1, 3) - points, p=2, dim=-1)
        idx = torch.argsort(dist, dim=-1, descending=False)[0]
        m = torch.zeros(len(idx))
        m[idx[:int(mask_ratio * len(idx))]] = 1
        out.append(m.bool())
    return torch.stack(out)
