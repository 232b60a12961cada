"""Input side: the oracle restatement of PointcloudScaleAndTranslate against the unmodified reference
class (authoring container), the product's RNG consumption against the oracle's (CPU), and the kernel against the oracle
bit for bit (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_model


def test_synthetic_clouds_generators_identical():
    from act_b200 import data
    assert torch.equal(data.synthetic_clouds(3, 1000, seed=5), ref_model.synthetic_clouds(3, 1000, seed=5))


def test_draw_consumes_numpy_rng_like_the_oracle():
    from act_b200 import data
    pc = torch.ones(7, 4, 3)
    np.random.seed(11)
    want = ref_model.scale_and_translate(pc.clone())
    np.random.seed(11)
    st = torch.from_numpy(data.PointcloudScaleAndTranslate().draw(7))
    got = pc * st[:, None, :3] + st[:, None, 3:]
    assert torch.equal(got, want)
    assert np.random.uniform() == np.random.RandomState(11).uniform(size=7 * 6 + 1)[-1]   # same stream position after


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="authoring container only")
def test_oracle_scale_and_translate_matches_real_reference():
    import importlib.util
    from oracle import shims
    shims.install()                                    # neutralises .cuda() on this GPU-less box
    spec = importlib.util.spec_from_file_location("ref_data_transforms", "/root/reference/datasets/data_transforms.py")
    data_transforms = importlib.util.module_from_spec(spec)        # the file alone: datasets/__init__ needs h5py
    spec.loader.exec_module(data_transforms)
    pc = ref_model.synthetic_clouds(5, 600, seed=3)
    np.random.seed(42)
    want = data_transforms.PointcloudScaleAndTranslate()(pc.clone())
    np.random.seed(42)
    got = ref_model.scale_and_translate(pc.clone())
    assert torch.equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(128, 1024), (3, 1000), (5, 601), (1, 2), (2, 8192)])
def test_scale_translate_kernel_bit_exact(shape):
    from act_b200 import data
    B, N = shape
    pc = ref_model.synthetic_clouds(B, N, seed=N)
    np.random.seed(B)
    want = ref_model.scale_and_translate(pc.clone())
    np.random.seed(B)
    dev = pc.cuda()
    got = data.PointcloudScaleAndTranslate()(dev)
    assert got.data_ptr() == dev.data_ptr()                       # in place, like the reference
    assert torch.equal(got.cpu(), want)


@pytest.mark.gpu
def test_scale_translate_extra_channels():
    from act_b200 import data
    pc = torch.randn(4, 100, 6)
    np.random.seed(1)
    want = ref_model.scale_and_translate(pc.clone())
    np.random.seed(1)
    got = data.PointcloudScaleAndTranslate()(pc.cuda())
    assert torch.equal(got.cpu(), want)


@pytest.mark.gpu
def test_shapenet_subsample_and_normalise_on_device():
    """data.ShapeNetOnDevice == ShapeNet.__getitem__ of the reference (random_sample on the carried-over permutation array
    + pc_norm), item by item, on the same numpy stream."""
    from act_b200 import data
    rng = np.random.default_rng(3)
    raw = (rng.standard_normal((6, 8192, 3)) * np.array([1.0, 0.5, 2.0]) + 0.3).astype(np.float32)
    np.random.seed(21)
    perm = np.arange(8192)
    want = []
    for i in range(6):                                   # ShapeNet55Dataset.py:45-63 verbatim semantics
        np.random.shuffle(perm)
        pc = raw[i][perm[:1024]]
        pc = pc - np.mean(pc, axis=0)
        want.append(pc / np.max(np.sqrt(np.sum(pc ** 2, axis=1))))
    np.random.seed(21)
    got = data.ShapeNetOnDevice(8192, 1024)(torch.from_numpy(raw).cuda())
    np.testing.assert_allclose(got.cpu().numpy(), np.stack(want), rtol=2e-5, atol=2e-6)
    assert abs(got.norm(dim=-1).max(dim=1)[0] - 1).max().item() < 1e-5


@pytest.mark.gpu
def test_device_side_random_mask():
    from act_b200 import ops
    seed = torch.tensor([5], dtype=torch.int64, device="cuda")
    for B, G, nm in [(128, 64, 38), (16, 512, 307), (3, 100, 0), (2, 64, 64)]:
        m = ops.mask_rand(seed, B, G, nm)
        assert m.dtype == torch.bool and m.shape == (B, G)
        assert torch.all(m.sum(1) == nm)
    a = ops.mask_rand(seed, 4096, 64, 38)
    assert torch.equal(a, ops.mask_rand(seed, 4096, 64, 38))
    assert not torch.equal(a, ops.mask_rand(torch.tensor([6], dtype=torch.int64, device="cuda"), 4096, 64, 38))
    freq = a.float().mean(0)                                    # every group masked with probability 38/64
    assert (freq - 38 / 64).abs().max().item() < 0.04           # sigma = 0.0077
    assert not torch.equal(a[0], a[1])
