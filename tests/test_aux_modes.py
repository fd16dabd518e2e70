"""`-p bilinear` and `-d` (reference network.rs:111-138, main.rs:139-141,153-155).  These two graphs have no
parameters and the reference commits no output for them, so they are checked for internal consistency (CPU)
and GPU-vs-oracle agreement (GPU); their semantics are unpinned and documented as such."""
import numpy as np
import pytest

from conftest import noise_image
from oracle import oracle as O


def test_oracle_aux_consistency():
    x = noise_image(12, 15, 3)
    assert np.allclose(O.bilinear(x, 1), x, atol=1e-6)           # factor 1 = identity (transfer functions invert)
    assert np.allclose(O.downsample(x, 1), x, atol=1e-6)
    c = np.full((6, 9, 3), 0.37, np.float32)
    assert np.allclose(O.bilinear(c, 3), 0.37, atol=1e-6) and np.allclose(O.downsample(c, 3), 0.37, atol=1e-6)
    d = O.downsample(x, 3)
    assert d.shape == (4, 5, 3)
    lin = np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4).astype(np.float64)
    want = lin[0:3, 3:6].mean(axis=(0, 1))
    want = np.where(want <= 0.0031308, want * 12.92, 1.055 * want ** (1 / 2.4) - 0.055)
    assert np.allclose(d[0, 1], want, atol=1e-6)                 # average pooling in linear light
    # bilinear upscaling then pooling by the same factor keeps flat regions and stays in range
    b = O.bilinear(x, 3)
    assert b.shape == (36, 45, 3) and b.min() >= x.min() - 1e-9 and b.max() <= x.max() + 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(12, 15), (1, 1), (40, 130), (7, 5)])
def test_gpu_aux_modes_match_oracle(shape):
    import rusty_sr_b200 as R
    H, W = shape
    x = noise_image(H, W, 17)
    assert np.abs(R.bilinear(x, 3) - O.bilinear(x, 3)).max() < 2e-5
    assert np.abs(R.downsample(x, 3) - O.downsample(x, 3)).max() < 2e-5
    img = (x * 255 + 0.5).astype(np.uint8)
    for fn, ofn in ((R.bilinear, O.bilinear), (R.downsample, O.downsample)):
        got = fn(img, 3)
        want = O.data_to_img(ofn(O.img_to_data(img), 3).astype(np.float32))
        assert got.shape == want.shape and (got[..., 3] == 255).all()
        assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
