"""The fused middle launch (rusty_sr_b200/csrc/conv_fused.cuh; rsr_create flag RSR_FLAG_FUSED_MID):
conv5|6|7 and conv8|9 (network.rs:65-70) in one kernel that keeps l2 in shared memory.  Same parity bar
as the split plan (tests/test_gpu_parity.py), plus agreement between the two plans."""
import numpy as np
import pytest

from conftest import Plans, load_png, noise_image, synthetic_kat_image
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import rusty_sr_b200 as R
    return R


@pytest.fixture(scope="module", params=["f16mx", "bf16x3"])
def pair(request, R, weights):
    pf = R.RSR_FLAG_BF16X3 if request.param == "bf16x3" else 0
    f = Plans({k: R.Upscaler(weights[k], flags=R.RSR_FLAG_FUSED_MID | pf) for k in weights}, request.param)
    s = Plans({k: R.Upscaler(weights[k], flags=R.RSR_FLAG_SPLIT_MID | pf) for k in weights}, request.param)
    yield f, s
    for d in (f, s):
        for u in d.values():
            u.close()


def errs(out, ref):
    return (float(np.abs(out.astype(np.float64) - ref).max()),
            float(np.abs(O.srgb_to_linear(out) - O.srgb_to_linear(ref)).max()))


def test_launch_plans(pair):
    f, s = pair
    x = noise_image(20, 40, 1)
    f["anime"].forward(x)
    s["anime"].forward(x)
    assert f["anime"].last_launch_count() == 4 and s["anime"].last_launch_count() == 5


@pytest.mark.parametrize("width,launches", [(100, 4), (124, 4), (125, 5), (128, 5), (248, 4), (256, 4), (260, 4), (372, 4), (384, 4), (1000, 4)])
def test_default_plan_is_chosen_by_width(R, weights, width, launches):
    """Without a plan flag the middle levels run fused unless the 124-pixel strips of the fused launch waste too much against
    the 128-pixel strips of the split launches: only W = 125..128 (2 strips vs 1) runs split; BASELINE config 3's 256-wide patches
    (3 strips vs 2) run fused (tools/plan_by_width.py)."""
    x = noise_image(6, width, width)
    with R.Upscaler(weights["anime"]) as u, R.Upscaler(weights["anime"], flags=R.RSR_FLAG_SPLIT_MID) as s:
        out = u.forward(x)
        assert u.last_launch_count() == launches, u.plan()
        assert np.abs(out - s.forward(x)).max() < 1e-4


@pytest.mark.parametrize("wname", ["imagenet", "imagenetlinear", "anime"])
@pytest.mark.parametrize("case", ["cartoon", "butterfly", "kat", "noise256"])
def test_fused_parity(pair, weights, wname, case):
    f, s = pair
    x = {"cartoon": lambda: O.img_to_data(load_png("cartoon_lr")),
         "butterfly": lambda: O.img_to_data(load_png("butterfly_lr")),
         "kat": lambda: synthetic_kat_image(),
         "noise256": lambda: noise_image(256, 256, 1234)}[case]()
    ref = O.forward(weights[wname], x, 3, np.float64)
    out = f[wname].forward(x)
    srgb, lin = errs(out, ref)
    tol = f.tol_noise if case == "noise256" else f.tol_natural
    assert lin < tol and srgb < tol, (srgb, lin)
    # the two plans differ only in fp32 summation order (fp16 + fp8 operands: an fp32-level difference in l2 can also flip the
    # rounding of an fp8 correction operand, worth ~2^-16 of the value)
    assert np.abs(out - s[wname].forward(x)).max() < (2e-5 if f.prec == "bf16x3" else 1e-4)


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (9, 1), (7, 5), (2, 123), (3, 124), (5, 125), (4, 126), (3, 127),
                                   (6, 128), (2, 247), (3, 248), (3, 249), (5, 373), (131, 4), (17, 257), (45, 300),
                                   (160, 130)])
def test_fused_odd_sizes(pair, weights, shape):
    """Sizes below / across the 124-pixel strips of the fused launch and the 128-pixel strips around it."""
    f, _ = pair
    H, W = shape
    x = noise_image(H, W, 100 + H * 1000 + W)
    ref = O.forward(weights["anime"], x, 3, np.float64)
    srgb, lin = errs(f["anime"].forward(x), ref)
    assert lin < f.tol_noise and srgb < f.tol_noise, (shape, srgb, lin)


def test_fused_golden_u8(pair):
    f, _ = pair
    lr, hr = load_png("cartoon_lr"), load_png("cartoon_rsa")
    img = f["anime"].forward_u8(lr)
    d = np.abs(img.astype(int) - hr.astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= f.tie_frac and (img[..., 3] == 255).all()


def test_fused_batch_and_bands_bit_identical(pair):
    """Work partition independence: batches and 7-row-halo bands reproduce the single frame exactly."""
    f, _ = pair
    u = f["imagenet"]
    xs = np.stack([noise_image(64, 140, 1234 + i) for i in range(5)])
    outs = u.forward(xs)
    for i in range(5):
        assert np.array_equal(outs[i], u.forward(xs[i]))
    x = noise_image(96, 150, 4321)
    full = u.forward(x)
    parts = []
    for b in range(4):
        y0, y1 = b * 96 // 4, (b + 1) * 96 // 4
        top, bot = min(7, y0), min(7, 96 - y1)
        parts.append(u.forward_band(x[y0 - top:y1 + bot], top, bot))
    assert np.array_equal(np.concatenate(parts, axis=0), full)


def test_fused_1080p(pair, weights):
    """BASELINE config 2 at full size: a 64-row slab against the oracle, the whole frame against the split plan."""
    f, s = pair
    x = noise_image(1080, 1920, 1234)
    out = f["imagenet"].forward(x)
    assert out.shape == (3240, 5760, 3) and np.isfinite(out).all()
    y0, y1 = 500, 564
    ref = O.forward(weights["imagenet"], x[y0 - 7:y1 + 7], 3, np.float64)[21:-21]
    srgb, lin = errs(out[3 * y0:3 * y1], ref)
    assert lin < f.tol_noise and srgb < f.tol_noise, (srgb, lin)
    assert np.abs(out - s["imagenet"].forward(x)).max() < (2e-5 if f.prec == "bf16x3" else 1e-4)
    # repeated calls are deterministic
    assert np.array_equal(out, f["imagenet"].forward(x))
