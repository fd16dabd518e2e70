"""Several GPUs from one host process (rsr_create_multi / rsr_multi_upscale_*; include/rsr_b200.h).
A device may be listed more than once, so the sharding logic (image i -> context i mod G; bands with 7-row halos, one
host thread per context) is exercised on a single-GPU box too.  (Named test_zz_*: the newest entry points run last.)"""
import numpy as np
import pytest

from conftest import TOL, noise_image
from oracle import oracle as O

gpu = pytest.mark.gpu


def test_no_gpu_or_bad_arguments_fail_loudly(weights):
    """Runs everywhere: without a GPU rsr_create_multi must fail with RSR_E_CUDA (never fall back); with one it must still
    reject an empty device list and a wrong parameter count with the reference's assert message (main.rs:162)."""
    import rusty_sr_b200 as R
    with pytest.raises(ValueError, match="Parameters selected do not have the size required"):
        R.MultiUpscaler(params=weights["imagenet"][:-1], devices=[0])
    with pytest.raises(R.RsrError):
        R.MultiUpscaler(params=weights["imagenet"], devices=[])
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if not have_gpu:
        with pytest.raises(R.RsrError) as e:
            R.MultiUpscaler(params=weights["imagenet"])
        assert e.value.code == -3   # RSR_E_CUDA


@gpu
@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
def test_multi_batch_equals_single_context(weights, devices):
    import rusty_sr_b200 as R
    xs = np.stack([noise_image(40, 150, 300 + i) for i in range(5)])
    with R.Upscaler(weights["imagenet"]) as u, R.MultiUpscaler(weights["imagenet"], devices=devices) as m:
        assert m.size() == len(devices)
        want = np.stack([u.forward(x) for x in xs])
        got = m.forward_batch(xs)
        assert np.array_equal(got, want)
        imgs = (xs * 255.0 + 0.5).astype(np.uint8)
        got8 = m.forward_batch_u8(imgs)
        assert np.array_equal(got8, np.stack([u.forward_u8(i) for i in imgs]))
    ref = O.forward(weights["imagenet"], xs[0], 3, np.float64)
    assert np.abs(got[0] - ref).max() < TOL["f16mx"]["noise"]


@gpu
@pytest.mark.parametrize("devices,shape", [([0, 0], (90, 150)), ([0, 0, 0], (31, 260)), ([0, 0, 0, 0], (3, 20)), ([0], (16, 16))])
def test_multi_frame_bands_bit_identical_to_the_whole_frame(weights, devices, shape):
    import rusty_sr_b200 as R
    x = noise_image(shape[0], shape[1], 4321)
    with R.Upscaler(weights["anime"]) as u, R.MultiUpscaler(weights["anime"], devices=devices) as m:
        assert np.array_equal(m.forward_frame(x), u.forward(x))
        img = (x * 255.0 + 0.5).astype(np.uint8)
        assert np.array_equal(m.forward_frame_u8(img), u.forward_u8(img))


@gpu
def test_multi_errors_name_the_device(weights):
    import rusty_sr_b200 as R
    with pytest.raises(R.RsrError):
        R.MultiUpscaler(weights["imagenet"], devices=[0, 9999])
    with R.MultiUpscaler(weights["imagenet"], devices=[0, 0]) as m:
        with pytest.raises(ValueError):
            m.forward_batch(np.zeros((2, 4, 4, 4), np.float32))
