"""The Stage-I dVAE training step on the act_b200 kernels (BASELINE config 3).

`DiscreteVAE` keeps the reference's class name, constructor contract (`cls(cfg.model)`), call surface
(`forward(inp, temperature=, hard=) -> 6-tuple`, `get_loss(ret, gt)`, `recon_loss`, `forward_tokenizer_features`) and
`state_dict` keys (/root/reference/models/dvae.py:278-357: encoder.*, dgcnn_1.*, codebook, dgcnn_2.*, decoder.*), so
tools/runner_autoencoder.py:137-146 drives it unchanged and its checkpoints load into the Stage-II teacher.

What runs where (forward AND backward):
    Group (FPS + kNN + centred gather)                      -> csrc/fps.cu, csrc/knn.cu                 (modules.Group)
    mini-PointNet Encoder, train-mode BatchNorm             -> layers.PointNetEncoderFn (tcgen05 GEMMs + pointnet.cu)
    DGCNN x2 (dvae.py:26-117): kNN k=4 among the centres    -> csrc/knn.cu;  every 1x1 conv as ONE token-level tcgen05
        GEMM per layer -- W.[x_k - x_q ; x_q] = Wa.x_k + (Wb - Wa).x_q, so the [B,2C,G,4] edge tensor is never built --
        with forward, dgrad and wgrad on layers.LinearFn; GroupNorm + LeakyReLU + max-over-k fused into one forward and
        two backward kernels per layer (csrc/dgcnn_train.cu, layers.DgcnnEdgeFn / GroupNormRowsFn)
    soft gumbel-softmax (dvae.py:346) + KL term (:320-332)  -> csrc/gumbel.cu (layers.GumbelSoftmaxFn / KlUniformFn: one
        pass per kernel, noise drawn in-kernel); the [BG,8192] x [8192,C] codebook einsum (dvae.py:347) on the tcgen05 GEMM
    FoldingNet Decoder (dvae.py:217-275)                    -> Linear / 1x1-conv layers on the tcgen05 GEMM; the K=5
        (seed, coarse-point) part of final_conv.0 is split off algebraically and its [BG,C] "global" part computed
        once per group instead of once per point; BatchNorm / ReLU are ATen ops
    ChamferDistanceL1 x2 (dvae.py:303-318)                  -> csrc/chamfer.cu forward + backward
CUDA only; there is no CPU path.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import layers, ops
from .models import register
from .modules import Encoder, Group
from .teacher import DGCNN


# ------------------------------------------------------------------------------------------ Chamfer losses
class ChamferDistanceL2(nn.Module):
    """extensions/chamfer_dist/__init__.py:28-45."""

    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def _filter(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1 = xyz1[torch.sum(xyz1, dim=2).ne(0)].unsqueeze(dim=0)
            xyz2 = xyz2[torch.sum(xyz2, dim=2).ne(0)].unsqueeze(dim=0)
        return xyz1, xyz2

    def forward(self, xyz1, xyz2):
        dist1, dist2 = ops.ChamferFunction.apply(*self._filter(xyz1, xyz2))
        return torch.mean(dist1) + torch.mean(dist2)


class ChamferDistanceL2_split(ChamferDistanceL2):
    """extensions/chamfer_dist/__init__.py:47-64."""

    def forward(self, xyz1, xyz2):
        dist1, dist2 = ops.ChamferFunction.apply(*self._filter(xyz1, xyz2))
        return torch.mean(dist1), torch.mean(dist2)


class ChamferDistanceL1(ChamferDistanceL2):
    """extensions/chamfer_dist/__init__.py:66-85."""

    def forward(self, xyz1, xyz2):
        dist1, dist2 = ops.ChamferFunction.apply(*self._filter(xyz1, xyz2))
        return (torch.mean(torch.sqrt(dist1)) + torch.mean(torch.sqrt(dist2))) / 2


# ------------------------------------------------------------------------------------------------- DGCNN
def dgcnn_forward(m, x, idx4, B, G):
    """DGCNN.forward (dvae.py:81-117) for x f32 [B*G, Cin], idx4 i64 [B,G,4] -> f32 [B, G, Cout]; differentiable.
    Per edge layer: one tcgen05 GEMM (P | Q) + one fused GroupNorm / LeakyReLU / max-over-k kernel, each with its backward."""
    f = layers.linear(x, m.input_trans.weight.squeeze(-1), m.input_trans.bias)            # [BG,128]
    feats = []
    for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
        conv, gn = layer[0], layer[1]
        pq = layers.EdgeLinearFn.apply(f, conv.weight.flatten(1))                          # [BG, 2*Cp] = (P | Q)
        f = layers.DgcnnEdgeFn.apply(pq, idx4, gn.weight, gn.bias, B, G, gn.eps, 0.2)      # [BG, Cp]
        feats.append(f)
    h5 = layers.linear(torch.cat(feats, dim=1), m.layer5[0].weight.squeeze(-1))            # [BG, Cout]
    gn = m.layer5[1]
    return layers.GroupNormRowsFn.apply(h5, gn.weight, gn.bias, B, G, gn.eps, 0.2).view(B, G, -1)


# ----------------------------------------------------------------------------------------------- Decoder
class Decoder(nn.Module):
    """FoldingNet decoder (dvae.py:217-275); same parameter slots (mlp.{0,2,4}, final_conv.{0,1,3,4,6})."""

    def __init__(self, encoder_channel, num_fine):
        super().__init__()
        self.num_fine = num_fine
        self.grid_size = 2
        self.num_coarse = self.num_fine // 4
        assert num_fine % 4 == 0
        self.mlp = nn.Sequential(nn.Linear(encoder_channel, 1024), nn.ReLU(inplace=True), nn.Linear(1024, 1024),
                                 nn.ReLU(inplace=True), nn.Linear(1024, 3 * self.num_coarse))
        self.final_conv = nn.Sequential(nn.Conv1d(encoder_channel + 3 + 2, 512, 1), nn.BatchNorm1d(512),
                                        nn.ReLU(inplace=True), nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512),
                                        nn.ReLU(inplace=True), nn.Conv1d(512, 3, 1))
        lin = torch.linspace(-0.05, 0.05, steps=self.grid_size, dtype=torch.float)
        a = lin.view(1, self.grid_size).expand(self.grid_size, self.grid_size).reshape(1, -1)
        b = lin.view(self.grid_size, 1).expand(self.grid_size, self.grid_size).reshape(1, -1)
        self.folding_seed = torch.cat([a, b], dim=0).view(1, 2, self.grid_size ** 2)       # 1 2 S (plain attribute)

    def forward(self, feature_global):
        bs, g, c = feature_global.shape
        BG, M, S, N = bs * g, self.num_coarse, self.grid_size ** 2, self.num_fine
        fg = feature_global.reshape(BG, c)
        h = F.relu(layers.linear(fg, self.mlp[0].weight, self.mlp[0].bias))
        h = F.relu(layers.linear(h, self.mlp[2].weight, self.mlp[2].bias))
        coarse = layers.linear(h, self.mlp[4].weight, self.mlp[4].bias).view(BG, M, 3)
        # final_conv.0 on cat([global (c), seed (2), coarse point (3)]) per fine point n = m*S + s  (dvae.py:259-266):
        #   W.[g; seed_s; p_m] + b = (Wg.g + b) [per group] + Ws.seed_s [per grid cell] + Wp.p_m [per coarse point]
        c0, bn0, _, c1, bn1, _, c2 = self.final_conv
        W0 = c0.weight.squeeze(-1)
        z_g = layers.linear(fg, W0[:, :c].contiguous(), c0.bias)                                        # [BG,512]
        if self.folding_seed.device != fg.device:          # a plain attribute in the reference: moved once, kept
            self.folding_seed = self.folding_seed.to(fg.device)
        seed = self.folding_seed[0].t().contiguous()                                       # [S,2]
        if self.training and M == 8 and S == 4:
            z = layers.FoldInputFn.apply(z_g, coarse, W0, seed)                            # csrc/folding.cu, act dtype
        else:
            z_s = seed @ W0[:, c:c + 2].t()                                                # [S,512]
            z_p = coarse @ W0[:, c + 2:].t()                                               # [BG,M,512]
            z = (z_g[:, None, None, :] + z_p[:, :, None, :] + z_s[None, None]).reshape(BG * N, 512)
        if self.training:      # BatchNorm1d + ReLU on the mini-PointNet's kernels, bf16 between the layers (layers.BnReluFn)
            a = layers.bn_relu(z, bn0)
            z = layers.linear(a, c1.weight.squeeze(-1), c1.bias, out_act=True)             # [BG*N,512]
            a = layers.bn_relu(z, bn1)
        else:
            a = F.relu(F.batch_norm(z, bn0.running_mean, bn0.running_var, bn0.weight, bn0.bias, False, bn0.momentum, bn0.eps))
            z = layers.linear(a, c1.weight.squeeze(-1), c1.bias)
            a = F.relu(F.batch_norm(z, bn1.running_mean, bn1.running_var, bn1.weight, bn1.bias, False, bn1.momentum, bn1.eps))
        # 512 -> 3: the output dimension is padded to 8 columns for the tensor-core tile
        W2 = torch.cat([c2.weight.squeeze(-1), c2.weight.new_zeros(5, 512)], dim=0)
        b2 = torch.cat([c2.bias, c2.bias.new_zeros(5)])
        off = layers.linear(a, W2, b2)[:, :3].view(BG, M, S, 3)
        fine = (off + coarse[:, :, None, :]).reshape(bs, g, N, 3)
        return coarse.view(bs, g, M, 3), fine


def gumbel_softmax(logits, tau, hard, gumbel=None):
    """torch.nn.functional.gumbel_softmax over the last dim (dvae.py:346); `gumbel` injects the noise for parity runs."""
    if gumbel is None:
        gumbel = -torch.empty_like(logits).exponential_().log()
    y = ((logits + gumbel) / tau).softmax(-1)
    if hard:
        one_hot = torch.zeros_like(y).scatter_(-1, y.argmax(-1, keepdim=True), 1.0)
        return one_hot - y.detach() + y
    return y


@register
class DiscreteVAE(nn.Module):
    def __init__(self, config, **kwargs):
        super().__init__()
        g = lambda k: config[k] if isinstance(config, dict) else getattr(config, k)   # noqa: E731
        self.group_size, self.num_group = g("group_size"), g("num_group")
        self.encoder_dims, self.tokens_dims = g("encoder_dims"), g("tokens_dims")
        self.decoder_dims, self.num_tokens = g("decoder_dims"), g("num_tokens")
        self.group_divider = Group(num_group=self.num_group, group_size=self.group_size)
        self.encoder = Encoder(encoder_channel=self.encoder_dims)
        self.dgcnn_1 = DGCNN(encoder_channel=self.encoder_dims, output_channel=self.num_tokens)
        self.codebook = nn.Parameter(torch.randn(self.num_tokens, self.tokens_dims))
        self.dgcnn_2 = DGCNN(encoder_channel=self.tokens_dims, output_channel=self.decoder_dims)
        self.decoder = Decoder(encoder_channel=self.decoder_dims, num_fine=self.group_size)
        self.build_loss_func()
        # in-kernel gumbel noise: a device int64 [1] the engine refreshes before every step (AutoencoderStep stages it next
        # to the schedules); None = one torch.randint draw on the device per forward
        self.gumbel_seed = None
        self._draws = 0

    def build_loss_func(self):
        self.loss_func_cdl1 = ChamferDistanceL1()
        self.loss_func_cdl2 = ChamferDistanceL2()

    def recon_loss(self, ret, gt):
        """dvae.py:303-318."""
        _, _, coarse, fine, group_gt, _ = ret
        bs, g = coarse.shape[:2]
        coarse = coarse.reshape(bs * g, -1, 3).contiguous()
        fine = fine.reshape(bs * g, -1, 3).contiguous()
        group_gt = group_gt.reshape(bs * g, -1, 3).contiguous()
        return self.loss_func_cdl1(coarse, group_gt) + self.loss_func_cdl1(fine, group_gt)

    def get_loss(self, ret, gt):
        """dvae.py:320-332: (reconstruction, KL(mean softmax || uniform))."""
        loss_recon = self.recon_loss(ret, gt)
        qbar = getattr(ret[-1], "_act_qbar", None)
        if qbar is not None:                  # mean softmax already formed by the fused forward (layers.GumbelSoftmaxFn)
            return loss_recon, layers.KlUniformFn.apply(qbar)
        log_qy = torch.log(F.softmax(ret[-1], dim=-1).mean(dim=1))
        log_uniform = torch.full_like(log_qy, math.log(1. / self.num_tokens))
        loss_klv = F.kl_div(log_qy, log_uniform, None, None, 'batchmean', log_target=True)
        return loss_recon, loss_klv

    def _features(self, neighborhood, center, temperature, hard, gumbel):
        B, G, _ = center.shape
        _, idx4, _ = ops.knn(center, center, 4, want_dist=False)                           # [B,G,4] i64, no grad
        tokens = self.encoder(neighborhood).reshape(B * G, -1)
        logits = dgcnn_forward(self.dgcnn_1, tokens, idx4, B, G)                           # [B,G,num_tokens]
        if self.training and not hard and self.num_tokens in ops.GUMBEL_V and logits.dtype == torch.float32:
            # fused soft gumbel-softmax (+ the mean softmax get_loss needs): csrc/gumbel.cu
            noise = None if gumbel is None else gumbel.reshape(B * G, -1).float().contiguous()
            seed = None
            if noise is None:
                seed = self.gumbel_seed if self.gumbel_seed is not None else torch.randint(
                    0, 2 ** 62, (1,), dtype=torch.int64, device=logits.device)
                self._draws += 1
            tau = temperature.reshape(1) if isinstance(temperature, torch.Tensor) else float(temperature)
            soft_one_hot, qbar = layers.GumbelSoftmaxFn.apply(logits.view(B * G, -1), tau, noise, seed,
                                                              self._draws & 0x7fffffff, B, G)
            logits._act_qbar = qbar          # get_loss(ret, gt) receives this very tensor object as ret[-1]
        else:
            soft_one_hot = gumbel_softmax(logits, temperature, hard, gumbel)
        # einsum('b g n, n c -> b g c') == Linear with weight codebook^T
        if soft_one_hot.dtype == ops.act_dtype() and self.tokens_dims % 8 == 0:
            sampled = layers.CodebookFn.apply(soft_one_hot.view(B * G, -1), self.codebook)
        else:
            sampled = layers.linear(soft_one_hot.view(B * G, -1), self.codebook.t().contiguous())
        feature = dgcnn_forward(self.dgcnn_2, sampled, idx4, B, G)
        return logits, feature

    def forward_tokenizer_features(self, neighborhood, center, return_global=True, gumbel=None):
        """dvae.py:334-341."""
        return self._features(neighborhood, center, 1., True, gumbel)[1]

    def forward(self, inp, temperature=1., hard=False, gumbel=None, **kwargs):
        """dvae.py:343-357."""
        neighborhood, center = self.group_divider(inp)
        logits, feature = self._features(neighborhood, center, temperature, hard, gumbel)
        coarse, fine = self.decoder(feature)
        with torch.no_grad():
            whole_fine = (fine + center.unsqueeze(2)).reshape(inp.size(0), -1, 3)
            whole_coarse = (coarse + center.unsqueeze(2)).reshape(inp.size(0), -1, 3)
        assert fine.size(2) == self.group_size
        return (whole_coarse, whole_fine, coarse, fine, neighborhood, logits)


def get_temp(niter, start=1.0, target=0.0625, ntime=100000):
    """tools/runner_autoencoder.py:43-53 with cfgs/autoencoder/pointbert_dvae.yaml:27-30."""
    if niter > ntime:
        return target
    return target + (start - target) * (1. + math.cos(math.pi * float(niter) / ntime)) / 2.


def get_kld_weight(niter, start=0.0, target=0.1, ntime=100000):
    """tools/runner_autoencoder.py:18-41 with cfgs/autoencoder/pointbert_dvae.yaml:33-36."""
    n = niter - 10000
    if n > ntime:
        return target
    if n < 0:
        return 0.
    return target + (start - target) * (1. + math.cos(math.pi * float(n) / ntime)) / 2.
