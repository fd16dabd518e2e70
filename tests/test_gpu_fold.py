"""Row-folded MMAs for conv0 / conv10 (conv_strip.cuh ConvCfg::FOLD; RSR_B200_FOLD=<launch bit mask>).  EXPERIMENTAL:
written after the round's GPU budget was spent, so never run on hardware yet — the schedule is checked on the CPU by
tests/test_fold_model.py, the kernels are off by default, and this module only runs when RSR_TEST_EXPERIMENTAL=1."""
import os

import numpy as np
import pytest

from conftest import noise_image
from oracle import oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("RSR_TEST_EXPERIMENTAL") != "1", reason="experimental kernels: set RSR_TEST_EXPERIMENTAL=1")]
TOL = 2e-4


def make(R, weights, mask):
    old = os.environ.get("RSR_B200_FOLD")
    os.environ["RSR_B200_FOLD"] = str(mask)
    try:
        return R.Upscaler(weights)
    finally:
        if old is None:
            del os.environ["RSR_B200_FOLD"]
        else:
            os.environ["RSR_B200_FOLD"] = old


@pytest.mark.parametrize("mask", [2, 32, 34])
@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (9, 130), (17, 257), (64, 300), (33, 1000)])
def test_fold_parity(weights, mask, shape):
    import rusty_sr_b200 as R
    H, W = shape
    x = noise_image(H, W, 11 + H * 1000 + W)
    ref = O.forward(weights["imagenet"], x, 3, np.float64)
    u, base = make(R, weights["imagenet"], mask), make(R, weights["imagenet"], 0)
    out, out0 = u.forward(x), base.forward(x)
    u.close()
    base.close()
    lin = float(np.abs(O.srgb_to_linear(out) - O.srgb_to_linear(ref)).max())
    assert lin < TOL and float(np.abs(out - ref).max()) < TOL
    assert np.abs(out - out0).max() < 2e-5   # same products, different summation order
