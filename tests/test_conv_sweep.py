"""BASELINE config 5: the generic 3x3 convolution on the sr_net kernel template, C in {32, 64, 128, 256}."""
import numpy as np
import pytest

from oracle import oracle as O


def make_case(H, W, ci, co, seed=99):
    rng = np.random.default_rng(seed)                      # N(0,1) * 1/sqrt(9C), seed 99
    x = rng.standard_normal((H, W, ci)).astype(np.float32)
    w = (rng.standard_normal((co, 3, 3, ci)) / np.sqrt(9 * ci)).astype(np.float32)
    return x, w


def test_oracle_conv2d_matches_direct_sum():
    x, w = make_case(5, 7, 32, 32, 1)
    out = O.conv2d(x, w)
    xp = np.pad(x.astype(np.float64), ((1, 1), (1, 1), (0, 0)))
    ref = np.zeros((5, 7, 32))
    for a in range(3):
        for b in range(3):
            ref += np.einsum("hwi,oi->hwo", xp[a:a + 5, b:b + 7], w[:, a, b].astype(np.float64))
    assert np.abs(out - ref).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("ci,co", [(32, 32), (64, 64), (128, 128), (256, 256), (64, 96), (32, 416), (96, 32)])
def test_conv2d_matches_oracle(ci, co):
    import rusty_sr_b200 as R
    H, W = 21, 140                                         # crosses a 128-pixel strip, odd row count
    x, w = make_case(H, W, ci, co)
    out, ms = R.conv2d_3x3(x, w)
    ref = O.conv2d(x, w)
    assert out.shape == ref.shape and ms > 0
    # outputs are ~N(0,1); the BF16x3 split keeps ~16 bits per operand
    assert np.abs(out - ref).max() < 2e-4, np.abs(out - ref).max()


@pytest.mark.gpu
def test_conv2d_rejects_unsupported_channels():
    import rusty_sr_b200 as R
    with pytest.raises(R.RsrError):
        R.conv2d_3x3(np.zeros((4, 4, 3), np.float32), np.zeros((32, 3, 3, 3), np.float32))
