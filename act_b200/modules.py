"""The reference's nn.Module surface for the hot path, re-implemented on the act_b200 kernels.

Same class names, constructor arguments, attribute names and therefore the same `state_dict()` keys and shapes
as /root/reference/models/dvae.py (Group, Encoder) and /root/reference/models/act.py (Mlp, Attention, Block,
TransformerEncoder, TransformerDecoder, VisableOnlyMaskTransformer) -- SURVEY.md App. A.5 -- so reference
checkpoints load unchanged.  Forward passes are NOT the reference's op-by-op PyTorch: each module routes to the
fused autograd Functions of layers.py (hand-written CUDA forward and backward).  CUDA only; there is no CPU path.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import layers, ops


class Group(nn.Module):
    """models/dvae.py:154-183: FPS centres + kNN neighbourhoods, centred.  Two kernel launches."""

    def __init__(self, num_group, group_size):
        super().__init__()
        self.num_group = num_group
        self.group_size = group_size

    def forward(self, xyz):
        neighborhood, center, idx, fps_idx = ops.group(xyz, self.num_group, self.group_size)
        self.last_idx, self.last_fps_idx = idx, fps_idx
        return neighborhood, center


class Encoder(nn.Module):
    """models/dvae.py:185-215 (mini-PointNet).  Parameters live in the same nn.Sequential slots as the
    reference (first_conv.{0,1,3}, second_conv.{0,1,3}); the Sequentials are never called."""

    def __init__(self, encoder_channel):
        super().__init__()
        self.encoder_channel = encoder_channel
        self.first_conv = nn.Sequential(nn.Conv1d(3, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True),
                                        nn.Conv1d(128, 256, 1))
        self.second_conv = nn.Sequential(nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True),
                                         nn.Conv1d(512, self.encoder_channel, 1))

    def forward(self, point_groups, n_keep=None):
        """point_groups [B,G,k,3] -> [B,G,C] (the reference contract).  With n_keep (groups flattened to [BG,k,3],
        the first n_keep of them wanted) -> [n_keep, C]: tokens of groups the caller discards are not computed."""
        c1, bn1, _, c2 = self.first_conv
        c3, bn2, _, c4 = self.second_conv
        bufs = (bn1.running_mean, bn1.running_var, bn1.num_batches_tracked, bn2.running_mean, bn2.running_var,
                bn2.num_batches_tracked)
        return layers.PointNetEncoderFn.apply(point_groups, self.training, bn1.momentum, bn1.eps, bufs, n_keep, c1.weight,
                                              c1.bias, bn1.weight, bn1.bias, c2.weight, c2.bias, c3.weight, c3.bias,
                                              bn2.weight, bn2.bias, c4.weight, c4.bias)


class Mlp(nn.Module):
    """models/act.py:25-42 (parameter container; computed inside the fused Block)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if drop != 0.:
            raise NotImplementedError("act_b200: dropout inside Mlp is 0 in every shipped config")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return layers.linear(layers.linear(x, self.fc1.weight, self.fc1.bias, gelu=True), self.fc2.weight,
                             self.fc2.bias)


class Attention(nn.Module):
    """models/act.py:45-69 (parameter container; computed inside the fused Block)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if qkv_bias or attn_drop != 0. or proj_drop != 0. or qk_scale is not None:
            raise NotImplementedError("act_b200: qkv_bias / dropout / qk_scale are unused by every shipped config")
        if dim // num_heads != 64:
            raise NotImplementedError("act_b200 attention kernels are specialised for head_dim 64")
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class Block(nn.Module):
    """models/act.py:72-90; also the duplicate at utils/transformer_layers.py:200-232 (same math, same keys)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.drop_path_rate = float(drop_path)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)

    def forward(self, x):
        return run_blocks([self], x, None, self.training)


def run_blocks(blocks, x, pos, training):
    rates = [b.drop_path_rate for b in blocks]
    gates = layers.drop_path_gates(rates, x.shape[0], x.device, training)
    return layers.transformer_stack(x, pos, blocks, blocks[0].attn.num_heads, blocks[0].norm1.eps, gates)


class TransformerEncoder(nn.Module):
    """models/act.py:93-112."""

    def __init__(self, embed_dim=768, depth=4, num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0.):
        super().__init__()
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate,
                  drop_path=drop_path_rate[i] if isinstance(drop_path_rate, list) else drop_path_rate)
            for i in range(depth)])

    def forward(self, x, pos):
        return run_blocks(list(self.blocks), x, pos, self.training)


class TransformerDecoder(nn.Module):
    """models/act.py:115-145 (xavier_uniform init of its Linears, LayerNorm on the last return_token_num)."""

    def __init__(self, embed_dim=384, depth=4, num_heads=6, mlp_ratio=4., qkv_bias=False, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1, norm_layer=nn.LayerNorm):
        super().__init__()
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate,
                  drop_path=drop_path_rate[i] if isinstance(drop_path_rate, list) else drop_path_rate)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Identity()
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, x, pos, return_token_num):
        x = run_blocks(list(self.blocks), x, pos, self.training)
        T = x.shape[1]
        return layers.layer_norm_rows(x, self.norm.weight, self.norm.bias, self.norm.eps, T - return_token_num,
                                      return_token_num)


def mask_center_rand(B, G, mask_ratio, device):
    """models/act.py:244-267: exactly int(mask_ratio*G) masked groups per cloud, host numpy RNG (kept for RNG
    parity with the reference), ONE H2D copy."""
    num_mask = int(mask_ratio * G)
    overall = np.zeros([B, G])
    for i in range(B):
        m = np.hstack([np.zeros(G - num_mask), np.ones(num_mask)])
        np.random.shuffle(m)
        overall[i, :] = m
    return torch.from_numpy(overall).to(torch.bool).to(device, non_blocking=True)


def pos_mlp(seq, x):
    """nn.Sequential(Linear(3,128), GELU, Linear(128,C)) (act.py:173-177, 1166-1170): K = 3 layer + GELU on CUDA cores, the
    128 -> C layer on the tcgen05 GEMM (layers.PosMlpFn, forward and backward)."""
    return layers.pos_mlp(seq, x)


class VisableOnlyMaskTransformer(nn.Module):
    """models/act.py:148-309: the MAE-style student encoder (mask, embed, keep visible tokens, 12 Blocks, LN)."""

    def __init__(self, config, **kwargs):
        super().__init__()
        self.config = config
        tc, dc = config.transformer_config, config.dvae_config
        self.mask_ratio = tc.mask_ratio
        self.embed_dim = tc.embed_dim
        self.cls_dim = tc.cls_dim
        self.depth = tc.depth
        self.drop_path_rate = tc.drop_path_rate
        self.num_heads = tc.num_heads
        self.encoder_dims = dc.encoder_dims
        self.encoder = Encoder(encoder_channel=self.encoder_dims)
        self.reduce_dim = (nn.Linear(self.encoder_dims, self.embed_dim) if self.encoder_dims != self.embed_dim
                           else nn.Identity())
        self.mask_type = tc.mask_type
        self.block_index = None                # mask_type 'block': staged random centre indices (see _mask_center_block)
        self.cls_token = nn.Parameter(torch.randn(1, 1, self.embed_dim))
        self.cls_pos = nn.Parameter(torch.randn(1, 1, self.embed_dim))
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, self.embed_dim))
        dpr = [x.item() for x in torch.linspace(0, self.drop_path_rate, self.depth)]
        self.blocks = TransformerEncoder(embed_dim=self.embed_dim, depth=self.depth, drop_path_rate=dpr,
                                         num_heads=self.num_heads)
        self.norm = nn.LayerNorm(self.embed_dim)
        self.num_tokens = dc.num_tokens
        self.lm_head = nn.Linear(self.embed_dim, self.num_tokens)
        self.cls_head = nn.Sequential(nn.Linear(self.embed_dim, self.cls_dim), nn.GELU(),
                                      nn.Linear(self.cls_dim, self.cls_dim))
        nn.init.trunc_normal_(self.cls_token, std=.02)
        nn.init.trunc_normal_(self.cls_pos, std=.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, (nn.Linear, nn.Conv1d)):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def _mask_center_rand(self, center, noaug=False):
        B, G, _ = center.shape
        if noaug or self.mask_ratio == 0:
            return torch.zeros(B, G, dtype=torch.bool, device=center.device)
        self.num_mask = int(self.mask_ratio * G)
        return mask_center_rand(B, G, self.mask_ratio, center.device)

    def _mask_center_block(self, center, noaug=False):
        """act.py:215-243: per cloud, the int(mask_ratio * G) centres nearest to a random centre.  The random centre index is
        drawn on the host with Python's `random.randint` exactly like the reference (same stream consumption), or read from
        `self.block_index` (device i32 [B]) when a caller -- engine.PretrainStep -- stages it for a captured step; distances,
        ranking and the mask are one kernel (csrc/tokens.cu act_mask_block), no host round trip."""
        import random
        B, G, _ = center.shape
        if noaug or self.mask_ratio == 0:
            return torch.zeros(B, G, dtype=torch.bool, device=center.device)
        index = self.block_index
        if index is None:
            index = torch.tensor([random.randint(0, G - 1) for _ in range(B)], dtype=torch.int32)
            index = index.pin_memory().to(center.device, non_blocking=True) if center.is_cuda else index
        return ops.mask_block(center, index, int(self.mask_ratio * G))

    def forward(self, neighborhood, center, only_cls_tokens=False, noaug=False, mask=None, return_extras=False):
        """-> (x_vis, mask) like the reference; with return_extras also the dict {order, centers_sorted, encoded} the Stage-II
        model consumes (returned, never kept on the module: a tensor with a grad_fn stored on a module would keep the previous
        step's autograd graph alive and release it in the middle of the next step -- e.g. inside a CUDA-graph capture)."""
        if self.mask_type not in ('rand', 'block'):
            raise NotImplementedError("act_b200: mask_type 'rand' (the shipped config) or 'block'")
        B, G, _ = center.shape
        if mask is None:
            mask = (self._mask_center_rand if self.mask_type == 'rand' else self._mask_center_block)(center, noaug=noaug)
        num_mask = 0 if (noaug or self.mask_ratio == 0) else int(self.mask_ratio * G)
        n_vis = G - num_mask
        # the permutation "visible groups first" (original order inside each part) replaces the reference's boolean
        # indexing tokens[~mask] / center[~mask] (act.py:281-284: nonzero + a host sync each); one kernel, no sync
        order = ops.mask_order(mask)
        if num_mask > 0 and isinstance(self.reduce_dim, nn.Identity):
            # The reference embeds all G groups and then throws the masked ones away (act.py:276-281).  Both
            # BatchNorms need every point, but the last conv + max-pool only matter for visible groups: reorder
            # the groups (all clouds' visible groups first) so that those rows are one contiguous prefix.
            nb_perm, centers_sorted, vis_center = ops.permute_groups(neighborhood, center, order, n_vis)
            x_vis = self.encoder(nb_perm, n_keep=B * n_vis)                       # [B*n_vis, C]
        else:
            tokens = self.encoder(neighborhood)
            if not isinstance(self.reduce_dim, nn.Identity):
                tokens = layers.linear(tokens, self.reduce_dim.weight, self.reduce_dim.bias)
            _, centers_sorted, vis_center = ops.permute_groups(None, center, order, n_vis, want_nb=False)
            x_vis = tokens if num_mask == 0 else torch.gather(tokens, 1, order[:, :n_vis, None].expand(-1, -1, tokens.shape[-1]))
        pos = pos_mlp(self.pos_embed, vis_center)                                  # [B*n_vis, C]
        x = layers.assemble_rows(x_vis, self.cls_token, B, n_vis, n_vis + 1, True)  # cat(cls_token, x_vis)
        pos = layers.assemble_rows(pos, self.cls_pos, B, n_vis, n_vis + 1, True)    # cat(cls_pos, pos)
        x = self.blocks(x, pos)
        x = layers.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        if only_cls_tokens:
            h = layers.linear(x[:, 0], self.cls_head[0].weight, self.cls_head[0].bias, gelu=True)
            return layers.linear(h, self.cls_head[2].weight, self.cls_head[2].bias)
        if return_extras:
            return x[:, 1:], mask, {"order": order, "centers_sorted": centers_sorted, "encoded": x}
        return x[:, 1:], mask
