"""Input side of the hot loop: the reference's GPU-side train transform on one kernel launch, and the
synthetic ShapeNet-shaped clouds the benchmark and tests feed (no dataset is available offline).

`PointcloudScaleAndTranslate` keeps the reference's class name, constructor arguments and call contract
(/root/reference/datasets/data_transforms.py:20-34: modifies `pc` in place and returns it) and consumes numpy's global
RNG in exactly the reference's order (per cloud: uniform(size=3) for the scale, then uniform(size=3) for the translation),
so a seeded run produces bit-identical batches.  The reference's loop does 2*B pageable H2D copies and 4*B launches per
step; this does one pinned async copy and one launch (csrc/augment.cu).
"""
import numpy as np
import torch

from . import ops


class PointcloudScaleAndTranslate(object):
    def __init__(self, scale_low=2. / 3., scale_high=3. / 2., translate_range=0.2):
        self.scale_low = scale_low
        self.scale_high = scale_high
        self.translate_range = translate_range

    def draw(self, bsize):
        """The reference's RNG consumption (data_transforms.py:28-30) -> float32 [B,6] (scale xyz | translate xyz)."""
        st = np.empty((bsize, 6), np.float32)
        for i in range(bsize):
            st[i, :3] = np.random.uniform(low=self.scale_low, high=self.scale_high, size=[3])     # float64 -> .float()
            st[i, 3:] = np.random.uniform(low=-self.translate_range, high=self.translate_range, size=[3])
        return st

    def __call__(self, pc):
        bsize = pc.size()[0]
        # fresh pinned staging per call (recycled by torch's host allocator only after the async copy has completed)
        st = torch.from_numpy(self.draw(bsize)).pin_memory().to(pc.device, non_blocking=True)
        if pc.shape[2] == 3 and pc.is_contiguous():
            return ops.scale_translate_(pc, st)
        xyz = pc[:, :, 0:3].contiguous()                      # clouds with extra channels: transform the xyz columns
        pc[:, :, 0:3] = ops.scale_translate_(xyz, st)
        return pc


def synthetic_clouds(B, N, seed=20231017):
    """SURVEY.md 8(d) synthetic ShapeNet-shaped clouds: ellipsoid-surface samples with jitter, centred and scaled into the
    unit sphere (datasets/ShapeNet55Dataset.py:45-67), then a per-cloud anisotropic scale U(2/3,3/2)^3 and translation
    U(-0.2,0.2)^3 (data_transforms.py:20-34).  CPU generator, fp32 [B,N,3] (tests/test_data.py checks it against the test infrastructure's own copy)."""
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


class ShapeNetOnDevice(object):
    """ShapeNet.__getitem__ (/root/reference/datasets/ShapeNet55Dataset.py:45-67) for a whole batch on the GPU: the loader
    workers only read the raw `.npy` clouds (8192 points); the per-item `random_sample` (np.random.shuffle of ONE persistent
    permutation array, first `npoints` entries kept -- the same numpy stream and the same carried-over array state as the
    reference) is drawn here on the host into one pinned [B, npoints] index buffer, and the gather + pc_norm (centroid,
    largest norm) run as one launch (csrc/augment.cu).  Results equal the reference's to fp32 summation order."""

    def __init__(self, n_raw=8192, npoints=1024):
        self.permutation = np.arange(n_raw)
        self.npoints = npoints

    def draw(self, bsize):
        sel = np.empty((bsize, self.npoints), np.int32)
        for i in range(bsize):
            np.random.shuffle(self.permutation)
            sel[i] = self.permutation[:self.npoints]
        return sel

    def __call__(self, raw):
        """raw: f32 [B, n_raw, 3] on the device (or pinned host memory) -> f32 [B, npoints, 3] on the device."""
        dev = raw.device if raw.is_cuda else torch.device("cuda", torch.cuda.current_device())
        sel = torch.from_numpy(self.draw(raw.shape[0])).pin_memory().to(dev, non_blocking=True)
        return ops.subsample_norm(raw.to(dev, non_blocking=True), sel)
