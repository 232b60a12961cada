"""Plain-PyTorch fp32 CPU restatement of the frozen Stage-I teacher's feature path.
TEST INFRASTRUCTURE ONLY (oracle): never imported by act_b200/.

Follows /root/reference/models/dvae.py: DGCNN (:26-117), ACTPromptedDiscreteVAEwithVIT.__init__ /
build_visual_embedding (:363-437), incorporate_prompt (:486-500), visual_embedding_deep_prompt (:536-576),
forward_tokenizer_features (:584-592).  The ViT blocks come from timm 0.5.4 (`vit_base_patch16_384`, requirements.txt:16;
not installed here, pretrained weights not obtainable offline): `VitBlock` restates timm's
vision_transformer.Block of that version (pre-LN, LayerNorm eps 1e-6, qkv WITH bias, exact GELU, no LayerScale).
PARITY PINNING: DGCNN, prompt handling, gumbel/codebook and the overall data flow are pinned against the
unmodified reference class run in this container through oracle/shims.py (tests/test_oracle_teacher.py and the golden
fixture tests/golden/teacher.npz); for that run timm.create_model is replaced by a stand-in built from VitBlock, so
the ViT block itself is "parity unpinned" against timm's binary -- it is the published ViT block.

State-dict keys equal the reference's (`encoder.*`, `dgcnn_1.*`, `codebook`, `dgcnn_2.*`, `decoder.*`,
`visual_embed.0.{i}.*`, `visual_embed.1.*`, `proj_pre`, `visual_pos_embed`, `proj_post`, `visual_prompt_token/pos`,
`deep_prompt_tokens/pos`), so a reference teacher checkpoint loads into it.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cpu_ref
from .ref_model import Encoder


class VitMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class VitAttention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class VitBlock(nn.Module):
    """timm 0.5.4 vision_transformer.Block (drop_path 0 in eval / for a frozen model)."""

    def __init__(self, dim=768, num_heads=12, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = VitAttention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = VitMlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class FakeTimmViT(nn.Module):
    """What `timm.create_model('vit_base_patch16_384')` must expose for dvae.py:405-411: .blocks, .norm, .embed_dim."""

    def __init__(self, dim=768, depth=12, num_heads=12):
        super().__init__()
        self.embed_dim = dim
        self.blocks = nn.Sequential(*[VitBlock(dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)


class DGCNN(nn.Module):
    """models/dvae.py:26-117 (k = 4 neighbours among the group centres, edge features [x_k - x_q ; x_q])."""

    def __init__(self, encoder_channel, output_channel):
        super().__init__()
        self.input_trans = nn.Conv1d(encoder_channel, 128, 1)

        def layer(cin, cout):
            return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, bias=False), nn.GroupNorm(4, cout),
                                 nn.LeakyReLU(negative_slope=0.2))
        self.layer1, self.layer2 = layer(256, 256), layer(512, 512)
        self.layer3, self.layer4 = layer(1024, 512), layer(1024, 1024)
        self.layer5 = nn.Sequential(nn.Conv1d(2304, output_channel, kernel_size=1, bias=False),
                                    nn.GroupNorm(4, output_channel), nn.LeakyReLU(negative_slope=0.2))

    @staticmethod
    def get_graph_feature(idx, x):
        """idx [B,G,4] (kNN of every centre among the centres), x [B,C,G] -> [B,2C,G,4] (dvae.py:59-79)."""
        B, C, G = x.shape
        xt = x.transpose(1, 2)                                            # B G C
        nb = torch.gather(xt[:, None].expand(-1, G, -1, -1), 2, idx[..., None].expand(-1, -1, -1, C))   # B G 4 C
        nb = nb.permute(0, 3, 1, 2)                                       # B C G 4
        xq = x[..., None].expand(-1, -1, -1, 4)
        return torch.cat((nb - xq, xq), dim=1)

    def forward(self, f, coor):
        with torch.no_grad():
            _, idx = cpu_ref.knn(coor.detach().numpy(), coor.detach().numpy(), 4)
            idx = torch.from_numpy(idx)                                   # B G 4, ascending by (distance, index)
        f = self.input_trans(f.transpose(1, 2))                           # B 128 G
        feats = []
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            f = layer(self.get_graph_feature(idx, f)).max(dim=-1)[0]
            feats.append(f)
        f = self.layer5(torch.cat(feats, dim=1))
        return f.transpose(-1, -2)


class FoldingDecoderParams(nn.Module):
    """Parameter container for the FoldingNet Decoder (dvae.py:217-244): only so that teacher checkpoints load
    strictly; it is not on the feature path."""

    def __init__(self, encoder_channel, num_fine):
        super().__init__()
        nc = num_fine // 4
        self.mlp = nn.Sequential(nn.Linear(encoder_channel, 1024), nn.ReLU(inplace=True), nn.Linear(1024, 1024),
                                 nn.ReLU(inplace=True), nn.Linear(1024, 3 * nc))
        self.final_conv = nn.Sequential(nn.Conv1d(encoder_channel + 5, 512, 1), nn.BatchNorm1d(512),
                                        nn.ReLU(inplace=True), nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512),
                                        nn.ReLU(inplace=True), nn.Conv1d(512, 3, 1))


class TeacherFeatures(nn.Module):
    """ACTPromptedDiscreteVAEwithVIT reduced to what ACT_PointDistillation.forward calls:
    forward_tokenizer_features(neighborhood, center, return_global=True)."""

    def __init__(self, encoder_dims=384, tokens_dims=384, decoder_dims=384, num_tokens=8192, group_size=32,
                 visual_embed_dim=768, vit_depth=12, vit_heads=12, num_prompt_token=64):
        super().__init__()
        self.num_prompt_token = num_prompt_token
        self.encoder = Encoder(encoder_dims)
        self.dgcnn_1 = DGCNN(encoder_dims, num_tokens)
        self.codebook = nn.Parameter(torch.randn(num_tokens, tokens_dims))
        self.dgcnn_2 = DGCNN(tokens_dims, decoder_dims)
        self.decoder = FoldingDecoderParams(decoder_dims, group_size)
        vit = FakeTimmViT(visual_embed_dim, vit_depth, vit_heads)
        self.visual_embed = nn.Sequential(vit.blocks, vit.norm)
        self.proj_pre = nn.Linear(tokens_dims, visual_embed_dim)
        self.visual_pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, visual_embed_dim))
        self.proj_post = nn.Linear(visual_embed_dim, tokens_dims)
        self.visual_prompt_token = nn.Parameter(torch.zeros(1, num_prompt_token, visual_embed_dim))
        self.visual_prompt_pos = nn.Parameter(torch.zeros(1, num_prompt_token, visual_embed_dim))
        self.deep_prompt_tokens = nn.Parameter(torch.zeros(vit_depth - 1, num_prompt_token, visual_embed_dim))
        self.deep_prompt_pos = nn.Parameter(torch.zeros(vit_depth - 1, num_prompt_token, visual_embed_dim))

    def _drop(self, t, keep):
        """prompt_dropout (p = 0.1, dvae.py:422,492,560): F.dropout in train mode, or an injected keep mask."""
        if keep is not None:
            return t * keep / 0.9
        return F.dropout(t, 0.1, self.training)

    def visual_embedding_deep_prompt(self, inp, center, keeps=None):
        B = inp.shape[0]
        P = self.num_prompt_token
        pos = self.visual_pos_embed(center)
        x = self.proj_pre(inp)
        x = torch.cat((self._drop(self.visual_prompt_token.expand(B, -1, -1), None if keeps is None else keeps[0]), x), 1)
        pos = torch.cat((self.visual_prompt_pos.expand(B, -1, -1), pos), dim=1)
        blocks = self.visual_embed[0]
        h = blocks[0](x + pos)
        for i in range(1, len(blocks)):
            if i <= self.deep_prompt_tokens.shape[0]:
                dp = self._drop(self.deep_prompt_tokens[i - 1].expand(B, -1, -1), None if keeps is None else keeps[i])
                h = torch.cat((dp, h[:, P:]), dim=1)
                pos = torch.cat((self.deep_prompt_pos[i - 1].expand(B, -1, -1), pos[:, P:]), dim=1)
            h = blocks[i](h + pos)
        return self.proj_post(self.visual_embed[-1](h)[:, P:])

    def forward_tokenizer_features(self, neighborhood, center, return_global=True, gumbel=None, keeps=None):
        logits = self.dgcnn_1(self.encoder(neighborhood), center)                # B G num_tokens
        if gumbel is None:
            one_hot = F.gumbel_softmax(logits, tau=1.0, dim=2, hard=True)
            sampled = torch.einsum('b g n, n c -> b g c', one_hot, self.codebook)
        else:       # hard gumbel-softmax forward value: one-hot of argmax(logits + gumbel noise)
            sampled = self.codebook[(logits + gumbel).argmax(-1)]
        feature = self.visual_embedding_deep_prompt(sampled, center, keeps)
        if return_global:
            feature = self.dgcnn_2(feature, center)
        return feature
