"""Factor-generic path (network.rs:37: the three `expand` convolutions have 3 * factor^2 output channels; main.rs:30-31: FACTOR is a
compile-time 3 "TODO expose as an option").  The bundled blobs are factor-3 nets, so factors 1, 2 and 4 run on synthetic parameters:
the shared layers of `imagenet.rsr` plus N(0, sigma) expand convolutions / bias (sigma = that of the bundled tensors, seed 1234),
checked against the CPU oracle, which implements sr_net for any factor."""
import numpy as np
import pytest

from conftest import TOL, noise_image, synthetic_params
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[1, 2, 4])
def net(request, weights):
    import rusty_sr_b200 as R
    f = request.param
    prm = synthetic_params(weights["imagenet"], f)
    assert prm.size == R.num_params(f) == R._lib.load().rsr_num_params(f)
    ups = {"f16mx": R.Upscaler(prm, factor=f), "bf16x3": R.Upscaler(prm, factor=f, flags=R.RSR_FLAG_BF16X3),
           "split": R.Upscaler(prm, factor=f, flags=R.RSR_FLAG_SPLIT_MID)}
    yield f, prm, ups
    for u in ups.values():
        u.close()


@pytest.mark.parametrize("shape", [(1, 1), (7, 5), (9, 130), (33, 257), (64, 300)])
def test_factor_parity(net, shape):
    f, prm, ups = net
    H, W = shape
    x = noise_image(H, W, 77 + H * 1000 + W)
    ref = O.forward(prm, x, f, np.float64)
    assert ref.shape == (H * f, W * f, 3)
    for name, u in ups.items():
        out = u.forward(x)
        assert out.shape == ref.shape
        tol = TOL["bf16x3" if name == "bf16x3" else "f16mx"]["noise"]
        lin = float(np.abs(O.srgb_to_linear(out) - O.srgb_to_linear(ref)).max())
        assert lin < tol and float(np.abs(out - ref).max()) < tol, (f, name, shape, lin)


def test_factor_u8_rgb_bands_and_batches(net):
    f, prm, ups = net
    u = ups["f16mx"]
    img = np.random.default_rng(5).integers(0, 256, size=(40, 150, 4), dtype=np.uint8)
    rgba = u.forward_u8(img)
    assert rgba.shape == (40 * f, 150 * f, 4) and (rgba[..., 3] == 255).all()
    assert np.array_equal(rgba, O.data_to_img(u.forward(O.img_to_data(img))))
    assert np.array_equal(u.forward_u8(img, rgb=True), rgba[..., :3])
    x = noise_image(60, 140, 4321)
    full = u.forward(x)
    parts = []
    for b in range(3):
        y0, y1 = b * 20, (b + 1) * 20
        top, bot = min(7, y0), min(7, 60 - y1)
        parts.append(u.forward_band(x[y0 - top:y1 + bot], top, bot))
    assert np.array_equal(np.concatenate(parts, axis=0), full)
    xs = np.stack([noise_image(30, 70, 900 + i) for i in range(3)])
    outs = u.forward(xs)
    for i in range(3):
        assert np.array_equal(outs[i], u.forward(xs[i]))


def test_factor4_rejects_the_fused_plan(weights):
    import rusty_sr_b200 as R
    with pytest.raises(R.RsrError) as e:
        R.Upscaler(synthetic_params(weights["imagenet"], 4), factor=4, flags=R.RSR_FLAG_FUSED_MID)
    assert e.value.code == R._lib.RSR_E_UNSUPPORTED
