"""Plain-PyTorch fp32 CPU restatement of the Stage-I dVAE training step (BASELINE config 3).
TEST INFRASTRUCTURE ONLY (oracle): never imported by act_b200/.

Follows /root/reference/models/dvae.py: Decoder (FoldingNet, :217-275), DiscreteVAE.__init__ (:278-296),
recon_loss (:303-318), get_loss (:320-332), forward (:343-357); the loss wrappers of
/root/reference/extensions/chamfer_dist/__init__.py:13-25 (ChamferFunction) and :67-85 (ChamferDistanceL1) over the
C oracle's chamfer (oracle/cpu_ref.c); the step of tools/runner_autoencoder.py:137-146
(`loss = loss_recon + kld_weight * loss_klv`).  Group / Encoder / DGCNN are the restatements of ref_model / ref_teacher.
PARITY PINNING: pinned against the UNMODIFIED reference `DiscreteVAE` imported through oracle/shims.py
(tests/test_oracle_dvae.py, golden fixture tests/golden/dvae_step.npz written by oracle/make_golden.py from the
reference itself).  State-dict keys equal the reference's (`encoder.*`, `dgcnn_1.*`, `codebook`, `dgcnn_2.*`,
`decoder.*`).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cpu_ref
from .ref_model import Encoder, Group
from .ref_teacher import DGCNN


class ChamferFn(torch.autograd.Function):
    """extensions/chamfer_dist/__init__.py:13-25 over the C restatement of chamfer.cu."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        d1, d2, i1, i2 = cpu_ref.chamfer_forward(xyz1.detach().contiguous().numpy(), xyz2.detach().contiguous().numpy())
        ctx.save_for_backward(xyz1, xyz2)
        ctx.idx = (i1, i2)
        return torch.from_numpy(d1), torch.from_numpy(d2)

    @staticmethod
    def backward(ctx, g1, g2):
        xyz1, xyz2 = ctx.saved_tensors
        gx1, gx2 = cpu_ref.chamfer_backward(xyz1.detach().contiguous().numpy(), xyz2.detach().contiguous().numpy(),
                                            ctx.idx[0], ctx.idx[1], g1.contiguous().numpy(), g2.contiguous().numpy())
        return torch.from_numpy(gx1), torch.from_numpy(gx2)


def chamfer_l1(xyz1, xyz2):
    """ChamferDistanceL1.forward (chamfer_dist/__init__.py:74-85)."""
    d1, d2 = ChamferFn.apply(xyz1, xyz2)
    return (torch.sqrt(d1).mean() + torch.sqrt(d2).mean()) / 2


def chamfer_l2(xyz1, xyz2):
    """ChamferDistanceL2.forward (chamfer_dist/__init__.py:35-45)."""
    d1, d2 = ChamferFn.apply(xyz1, xyz2)
    return d1.mean() + d2.mean()


class Decoder(nn.Module):
    """FoldingNet decoder, dvae.py:217-275."""

    def __init__(self, encoder_channel, num_fine):
        super().__init__()
        self.num_fine, self.grid_size, self.num_coarse = num_fine, 2, num_fine // 4
        self.mlp = nn.Sequential(nn.Linear(encoder_channel, 1024), nn.ReLU(inplace=True), nn.Linear(1024, 1024),
                                 nn.ReLU(inplace=True), nn.Linear(1024, 3 * self.num_coarse))
        self.final_conv = nn.Sequential(nn.Conv1d(encoder_channel + 3 + 2, 512, 1), nn.BatchNorm1d(512),
                                        nn.ReLU(inplace=True), nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512),
                                        nn.ReLU(inplace=True), nn.Conv1d(512, 3, 1))
        lin = torch.linspace(-0.05, 0.05, steps=2, dtype=torch.float)
        a = lin.view(1, 2).expand(2, 2).reshape(1, -1)
        b = lin.view(2, 1).expand(2, 2).reshape(1, -1)
        self.folding_seed = torch.cat([a, b], dim=0).view(1, 2, 4)          # plain attribute, as in the reference

    def forward(self, feature_global):
        bs, g, c = feature_global.shape
        fg = feature_global.reshape(bs * g, c)
        coarse = self.mlp(fg).reshape(bs * g, self.num_coarse, 3)
        rep = coarse.unsqueeze(2).expand(-1, -1, 4, -1).reshape(bs * g, self.num_fine, 3).transpose(2, 1)   # BG 3 N
        seed = self.folding_seed.unsqueeze(2).expand(bs * g, -1, self.num_coarse, -1).reshape(bs * g, -1, self.num_fine)
        feat = torch.cat([fg.unsqueeze(2).expand(-1, -1, self.num_fine), seed, rep], dim=1)
        fine = self.final_conv(feat) + rep
        fine = fine.reshape(bs, g, 3, self.num_fine).transpose(-1, -2)
        return coarse.reshape(bs, g, self.num_coarse, 3), fine


class DiscreteVAE(nn.Module):
    """dvae.py:278-357.  `gumbel` (optional [B,G,num_tokens]) injects the noise of F.gumbel_softmax for parity runs."""

    def __init__(self, group_size=32, num_group=64, encoder_dims=256, tokens_dims=256, decoder_dims=256,
                 num_tokens=8192):
        super().__init__()
        self.group_size, self.num_group, self.num_tokens = group_size, num_group, num_tokens
        self.group_divider = Group(num_group, group_size)
        self.encoder = Encoder(encoder_dims)
        self.dgcnn_1 = DGCNN(encoder_dims, num_tokens)
        self.codebook = nn.Parameter(torch.randn(num_tokens, tokens_dims))
        self.dgcnn_2 = DGCNN(tokens_dims, decoder_dims)
        self.decoder = Decoder(decoder_dims, group_size)

    def forward(self, inp, temperature=1., hard=False, gumbel=None, **kwargs):
        neighborhood, center = self.group_divider(inp)
        logits = self.dgcnn_1(self.encoder(neighborhood), center)
        if gumbel is None:
            soft_one_hot = F.gumbel_softmax(logits, tau=temperature, dim=2, hard=hard)
        else:
            soft_one_hot = gumbel_softmax_with_noise(logits, gumbel, temperature, hard)
        sampled = torch.einsum('b g n, n c -> b g c', soft_one_hot, self.codebook)
        feature = self.dgcnn_2(sampled, center)
        coarse, fine = self.decoder(feature)
        with torch.no_grad():
            whole_fine = (fine + center.unsqueeze(2)).reshape(inp.size(0), -1, 3)
            whole_coarse = (coarse + center.unsqueeze(2)).reshape(inp.size(0), -1, 3)
        return (whole_coarse, whole_fine, coarse, fine, neighborhood, logits)

    def recon_loss(self, ret, gt):
        _, _, coarse, fine, group_gt, _ = ret
        bs, g = coarse.shape[:2]
        coarse = coarse.reshape(bs * g, -1, 3).contiguous()
        fine = fine.reshape(bs * g, -1, 3).contiguous()
        group_gt = group_gt.reshape(bs * g, -1, 3).contiguous()
        return chamfer_l1(coarse, group_gt) + chamfer_l1(fine, group_gt)

    def get_loss(self, ret, gt):
        loss_recon = self.recon_loss(ret, gt)
        log_qy = torch.log(F.softmax(ret[-1], dim=-1).mean(dim=1))
        log_uniform = torch.full_like(log_qy, math.log(1. / self.num_tokens))
        loss_klv = F.kl_div(log_qy, log_uniform, None, None, 'batchmean', log_target=True)
        return loss_recon, loss_klv


def gumbel_softmax_with_noise(logits, gumbel, tau, hard):
    """torch.nn.functional.gumbel_softmax with its noise made explicit: softmax((logits + g) / tau) over the last dim,
    straight-through one-hot when hard."""
    y = ((logits + gumbel) / tau).softmax(-1)
    if hard:
        one_hot = torch.zeros_like(y).scatter_(-1, y.argmax(-1, keepdim=True), 1.0)
        return one_hot - y.detach() + y
    return y


def temperature_schedule(step, start=1.0, target=0.0625, ntime=100000):
    """tools/runner_autoencoder.py:43-53 (get_temp): cosine from start to target over ntime iterations."""
    if step > ntime:
        return target
    return target + (start - target) * (1. + math.cos(math.pi * float(step) / ntime)) / 2.


def kld_weight_schedule(step, start=0.0, target=0.1, ntime=100000):
    """tools/runner_autoencoder.py:18-41 (compute_loss): 0 for the first 10 000 iterations, then the same cosine form
    over `ntime` iterations (cfgs/autoencoder/pointbert_dvae.yaml:33-38)."""
    n = step - 10000
    if n > ntime:
        return target
    if n < 0:
        return 0.
    return target + (start - target) * (1. + math.cos(math.pi * float(n) / ntime)) / 2.
