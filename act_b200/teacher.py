"""The frozen Stage-I teacher's feature path on the act_b200 kernels.

`ACTPromptedDiscreteVAEwithVIT` keeps the reference's class name, constructor contract (`cls(dvae_config)`) and
`state_dict` keys (/root/reference/models/dvae.py:360-437: encoder.*, dgcnn_1.*, codebook, dgcnn_2.*, decoder.*,
visual_embed.0.{i}.*, visual_embed.1.*, proj_pre, visual_pos_embed, proj_post, visual_prompt_token/pos,
deep_prompt_tokens/pos), so a reference teacher checkpoint loads unchanged; only
`forward_tokenizer_features` (dvae.py:584-592) -- the call ACT_PointDistillation.forward makes under no_grad
(act.py:1216-1217) -- is implemented, forward only:

    mini-PointNet (train-mode BatchNorm, as the reference runs it)  -> layers.PointNetEncoderFn
    DGCNN x2 (dvae.py:26-117)                                        -> one tcgen05 GEMM + one fused kernel per layer
    hard gumbel-softmax + codebook (dvae.py:587-588)                 -> fused GroupNorm/LeakyReLU/+noise/arg-max, gather
    VPT-deep prompted ViT-B, 12 blocks, 64 prompts (keys/values only) + 64 tokens -> tcgen05 GEMMs, LayerNorm, mma.sync attention

The pretrained ViT / dVAE weights are not obtainable offline; the module is exercised with deterministic weights.
"""
import torch
import torch.nn as nn

from . import layers, ops
from .modules import Encoder


class _VitMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)


class _VitAttention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class VitBlock(nn.Module):
    """Parameter container with timm 0.5.4's Block layout (norm1, attn.qkv/proj, norm2, mlp.fc1/fc2; LN eps 1e-6)."""

    def __init__(self, dim=768, num_heads=12, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _VitAttention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _VitMlp(dim, int(dim * mlp_ratio))


class DGCNN(nn.Module):
    """Parameter container of models/dvae.py:26-57."""

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


class _FoldingDecoderParams(nn.Module):
    """FoldingNet Decoder parameters (dvae.py:217-244): present so teacher checkpoints load strictly; not on this path."""

    def __init__(self, encoder_channel, num_fine):
        super().__init__()
        nc = num_fine // 4
        self.mlp = nn.Sequential(nn.Linear(encoder_channel, 1024), nn.ReLU(inplace=True), nn.Linear(1024, 1024),
                                 nn.ReLU(inplace=True), nn.Linear(1024, 3 * nc))
        self.final_conv = nn.Sequential(nn.Conv1d(encoder_channel + 5, 512, 1), nn.BatchNorm1d(512),
                                        nn.ReLU(inplace=True), nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512),
                                        nn.ReLU(inplace=True), nn.Conv1d(512, 3, 1))


def _bf(t):
    return t.detach().to(torch.bfloat16).contiguous()


class ACTPromptedDiscreteVAEwithVIT(nn.Module):
    def __init__(self, config, **kwargs):
        super().__init__()
        g = lambda k, d: getattr(config, k, d) if not isinstance(config, dict) else config.get(k, d)  # noqa: E731
        self.group_size, self.num_group = g("group_size", 32), g("num_group", 64)
        self.encoder_dims, self.tokens_dims = g("encoder_dims", 384), g("tokens_dims", 384)
        self.decoder_dims, self.num_tokens = g("decoder_dims", 384), g("num_tokens", 8192)
        self.visual_embed_dim = g("visual_embed_dim", 768)
        self.num_prompt_token = g("num_prompt_token", 64)
        depth, heads = g("visual_embed_depth", 12), g("visual_embed_heads", 12)
        if not g("use_deep_prompt", True) or self.num_prompt_token <= 0:
            raise NotImplementedError("act_b200 teacher: the shipped config (VPT-deep, 64 prompt tokens) only")
        self.encoder = Encoder(encoder_channel=self.encoder_dims)
        self.dgcnn_1 = DGCNN(self.encoder_dims, self.num_tokens)
        self.codebook = nn.Parameter(torch.randn(self.num_tokens, self.tokens_dims))
        self.dgcnn_2 = DGCNN(self.tokens_dims, self.decoder_dims)
        self.decoder = _FoldingDecoderParams(self.decoder_dims, self.group_size)
        D = self.visual_embed_dim
        self.visual_embed = nn.Sequential(nn.Sequential(*[VitBlock(D, heads) for _ in range(depth)]),
                                          nn.LayerNorm(D, eps=1e-6))
        self.visual_embed_depth = depth
        self.proj_pre = nn.Linear(self.tokens_dims, D)
        self.visual_pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, D))
        self.proj_post = nn.Linear(D, self.tokens_dims)
        P = self.num_prompt_token
        self.visual_prompt_token = nn.Parameter(torch.zeros(1, P, D))
        self.visual_prompt_pos = nn.Parameter(torch.randn(1, P, D))
        self.deep_prompt_tokens = nn.Parameter(torch.zeros(depth - 1, P, D))
        self.deep_prompt_pos = nn.Parameter(torch.randn(depth - 1, P, D))
        for t in (self.visual_prompt_token, self.visual_prompt_pos, self.deep_prompt_tokens, self.deep_prompt_pos):
            nn.init.trunc_normal_(t, std=.02)
        self._cache = None
        # optional device int64 [1] the caller refreshes before every call (engine.PretrainStep's pipelined mode stages it
        # from the host): the per-call seed of the in-kernel gumbel / prompt-dropout draws is then read from it instead of
        # being drawn from torch's default CUDA generator, whose state a CONCURRENTLY replaying graph also advances
        self.seed_buffer = None
        import os
        self.sm_cap = int(os.environ.get("ACT_B200_TEACHER_SM_CAP", "0"))

    # ---- frozen bf16 operand cache (weights never change: built once, dropped on load / device move) ----------
    def _apply(self, fn, *a, **k):
        self._cache = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._cache = None
        return super().load_state_dict(*a, **k)

    def _dgcnn_cache(self, m):
        c = {"it_w": _bf(m.input_trans.weight.squeeze(-1)), "it_b": m.input_trans.bias.detach().float()}
        for i, layer in enumerate((m.layer1, m.layer2, m.layer3, m.layer4)):
            W = layer[0].weight.detach().float().flatten(1)               # [Cp, 2*Cin]
            cin = W.shape[1] // 2
            Wa, Wb = W[:, :cin], W[:, cin:]
            c[f"w{i}"] = _bf(torch.cat([Wa, Wb - Wa], dim=0))             # [2*Cp, Cin]: P = x Wa^T | Q = x (Wb-Wa)^T
            c[f"g{i}"], c[f"b{i}"] = layer[1].weight.detach().float(), layer[1].bias.detach().float()
        c["w5"] = _bf(m.layer5[0].weight.squeeze(-1))
        c["g5"], c["b5"] = m.layer5[1].weight.detach().float(), m.layer5[1].bias.detach().float()
        return c

    def _prepare(self):
        if self._cache is not None:
            return self._cache
        c = {"d1": self._dgcnn_cache(self.dgcnn_1), "d2": self._dgcnn_cache(self.dgcnn_2),
             "pre_w": _bf(self.proj_pre.weight), "post_w": _bf(self.proj_post.weight),
             "pos2_w": _bf(self.visual_pos_embed[2].weight), "codebook": _bf(self.codebook), "blocks": []}
        for blk in self.visual_embed[0]:
            c["blocks"].append({"qkv": _bf(blk.attn.qkv.weight), "proj": _bf(blk.attn.proj.weight),
                                "fc1": _bf(blk.mlp.fc1.weight), "fc2": _bf(blk.mlp.fc2.weight)})
        for p in self.encoder.parameters():          # frozen: layers.shadow() finds the bf16 copy instead of casting per call
            if p.dim() > 1:
                p._act_shadow = _bf(p)
        self._cache = c
        return c

    # ---- pieces ----------------------------------------------------------------------------------------------
    def _dgcnn(self, m, c, x, idx4, B, G, noise=None, seed=None, want_labels=False):
        """x f32 [B*G, Cin] -> layer5 activations f32 [B*G, Cout], or (noise given) arg-max labels i32 [B*G]."""
        dev = x.device
        xb = x if x.dtype == torch.bfloat16 else ops.cast_rows(x.contiguous())              # C % 128 == 0
        f = ops.gemm(xb, c["it_w"], bias=c["it_b"])                                         # [BG,128]
        feats = torch.empty(B * G, 2304, dtype=torch.bfloat16, device=dev)
        off = 0
        for i, layer in enumerate((m.layer1, m.layer2, m.layer3, m.layer4)):
            Cp = layer[0].weight.shape[0]
            pq = ops.gemm(f, c[f"w{i}"], out_dtype=torch.float32)                           # [BG, 2*Cp]
            f = feats[:, off:off + Cp]
            ops.dgcnn_edge_gn(pq, idx4, c[f"g{i}"], c[f"b{i}"], B, G, Cp, layer[1].eps, 0.2, f)
            off += Cp
        h5 = ops.gemm(feats, c["w5"])                                                       # [BG, Cout] bf16
        if not want_labels:
            return ops.gn_rows(h5, c["g5"], c["b5"], B, G, m.layer5[1].eps, 0.2)
        return ops.gn_rows(h5, c["g5"], c["b5"], B, G, m.layer5[1].eps, 0.2, noise=noise, seed=seed)

    def _vit_block(self, x, pos_tok, tok, ppos, blk, w, B, G, keep, seed, draw_id):
        """One prompted block on the G token rows (the prompt rows' block output is dead in the reference -- the next
        block overwrites them, dvae.py:556-566 -- so prompts only supply keys / values): fused (prompt rebuild + pos add +
        norm1) -> q,k,v of the tokens / k,v of the prompts -> prefix attention -> proj(+resid) -> norm2 -> fc1(GELU) ->
        fc2(+resid).  x, returned value: f32 [B*G, D]."""
        H = blk.attn.num_heads
        eps = blk.norm1.eps
        P, D = self.num_prompt_token, self.visual_embed_dim
        p_drop = 0.1 if (self.training or keep is not None) else 0.0
        xs, h_tok, h_prm = ops.vit_ln1_fwd(x, pos_tok, tok, ppos, blk.norm1.weight, blk.norm1.bias, eps, B, G, P,
                                           keep=keep, seed=seed, draw_id=draw_id, p_drop=p_drop)
        qkv_t = ops.gemm(h_tok, w["qkv"], bias=blk.attn.qkv.bias)
        kv_p = ops.gemm(h_prm, w["qkv"][D:], bias=blk.attn.qkv.bias[D:])
        o = ops.attention_prefix_fwd(qkv_t, kv_p, B, G, P, H, (D // H) ** -0.5)
        xmid = ops.gemm(o, w["proj"], bias=blk.attn.proj.bias, resid=xs, out_dtype=torch.float32)
        h2, _, _, _ = ops.layernorm_fwd(xmid, blk.norm2.weight, blk.norm2.bias, eps, save_stats=False)
        a = ops.gemm(h2, w["fc1"], bias=blk.mlp.fc1.bias, act=ops.ACT_GELU)
        return ops.gemm(a, w["fc2"], bias=blk.mlp.fc2.bias, resid=xmid, out_dtype=torch.float32)

    def _visual(self, c, sampled, center, B, G, keeps, seed=None, out_dtype=torch.float32):
        """visual_embedding_deep_prompt (dvae.py:536-576)."""
        pe = self.visual_pos_embed
        pos_tok = ops.gemm(ops.pos_mlp1_fwd(center.reshape(B * G, 3), pe[0].weight, pe[0].bias), c["pos2_w"],
                           bias=pe[2].bias, out_dtype=torch.float32)
        if sampled.dtype != torch.bfloat16:
            sampled = ops.cast_rows(sampled.float().contiguous())
        x = ops.gemm(sampled, c["pre_w"], bias=self.proj_pre.bias, out_dtype=torch.float32)
        blocks = self.visual_embed[0]
        for i, blk in enumerate(blocks):
            if i == 0:
                tok, ppos = self.visual_prompt_token[0], self.visual_prompt_pos[0]
            elif i <= self.deep_prompt_tokens.shape[0]:
                tok, ppos = self.deep_prompt_tokens[i - 1], self.deep_prompt_pos[i - 1]
            else:
                raise NotImplementedError("act_b200 teacher: one deep prompt per block after the first (the shipped config)")
            keep = None if keeps is None else keeps[i].float().contiguous()
            x = self._vit_block(x, pos_tok, tok.detach(), ppos.detach(), blk, c["blocks"][i], B, G, keep, seed, i)
        norm = self.visual_embed[1]
        y, _, _, _ = ops.layernorm_fwd(x, norm.weight, norm.bias, norm.eps, save_stats=False)
        return ops.gemm(y, c["post_w"], bias=self.proj_post.bias, out_dtype=out_dtype)      # [BG, tokens_dims]

    @torch.no_grad()
    def forward_tokenizer_features(self, neighborhood, center, return_global=True, gumbel=None, keeps=None):
        """dvae.py:584-592.  gumbel (optional f32 [B,G,num_tokens]) / keeps (optional list of [B,P,D] 0/1 masks)
        inject the two random draws (gumbel noise, prompt dropout) for parity runs; default: drawn here."""
        # the frozen teacher always runs the bf16 speed mode (its own parity bounds: tests/test_gpu_teacher.py); sm_cap:
        # see ops.gemm_sm_cap (0 = its GEMMs may take every SM)
        with ops.precision("bf16"), ops.gemm_sm_cap(self.sm_cap):
            return self._features(neighborhood, center, return_global, gumbel, keeps)

    def _features(self, neighborhood, center, return_global, gumbel, keeps):
        c = self._prepare()
        B, G, _ = center.shape
        tokens = self.encoder(neighborhood).reshape(B * G, -1)
        _, idx4, _ = ops.knn(center, center, 4, want_dist=False)                            # [B,G,4] i64
        # one graph-safe draw per call (torch's Philox state advances on every CUDA-graph replay); the kernels derive
        # the gumbel noise and the prompt-dropout masks from it on the fly
        seed = self.seed_buffer
        if seed is None:
            seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=center.device)
        if gumbel is not None:
            gumbel = gumbel.reshape(B * G, self.num_tokens).float().contiguous()
        labels = self._dgcnn(self.dgcnn_1, c["d1"], tokens, idx4, B, G, noise=gumbel,
                             seed=seed if gumbel is None else None, want_labels=True)
        self.last_labels = labels
        sampled = ops.embedding_bf16(c["codebook"], labels)                                 # one-hot @ codebook, bf16
        # the ViT's output feeds only dgcnn_2's first GEMM when return_global: emit its bf16 operand directly
        feature = self._visual(c, sampled, center.float(), B, G, keeps, seed,
                               out_dtype=torch.bfloat16 if return_global else torch.float32)
        if return_global:
            feature = self._dgcnn(self.dgcnn_2, c["d2"], feature, idx4, B, G)
        return feature.view(B, G, -1)

    def forward(self, *a, **k):
        raise NotImplementedError("act_b200 teacher implements forward_tokenizer_features only (the Stage-II call); "
                                  "the Stage-I training forward is SURVEY row f2")
