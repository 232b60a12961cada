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
