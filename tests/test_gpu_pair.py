"""CTA-pair kernels (rusty_sr_b200/csrc/conv_pair.cuh, conv_fused.cuh: tcgen05 cta_group::2, two SMs per MMA): conv1|2|3 and
the fused middle launch run on pairs by default.  Every pair mask (test hook rsr_debug_create: bit 2 = conv1|2|3, bit 3 = the
fused middle launch) must meet the same parity bar and agree with the single-CTA kernels, in both operand formats."""
import numpy as np
import pytest

from conftest import TOL, load_png, noise_image
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["f16mx", "bf16x3"])
def plans(request, weights):
    import rusty_sr_b200 as R
    pf = R.RSR_FLAG_BF16X3 if request.param == "bf16x3" else 0
    mk = lambda mask, flags=0: R.Upscaler(weights["imagenet"], flags=flags | pf, pair_mask=mask)
    d = {m: mk(m, R.RSR_FLAG_FUSED_MID) for m in (0, 4, 8, 12)}
    d["nopairs-flag"] = mk(12, R.RSR_FLAG_FUSED_MID | R.RSR_FLAG_NO_PAIRS)
    d["split-0"] = mk(0, R.RSR_FLAG_SPLIT_MID)
    d["split-4"] = mk(4, R.RSR_FLAG_SPLIT_MID)
    d["tol"] = TOL[request.param]["noise"]
    d["flags"] = pf
    yield d
    for u in d.values():
        if hasattr(u, "close"):
            u.close()


@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (5, 130), (9, 257), (64, 300), (171, 256), (33, 1000)])
def test_pair_plans_parity_and_agreement(plans, weights, shape):
    H, W = shape
    x = noise_image(H, W, 7 + H * 1000 + W)
    ref = O.forward(weights["imagenet"], x, 3, np.float64)
    outs = {m: u.forward(x) for m, u in plans.items() if hasattr(u, "forward")}
    for m, out in outs.items():
        lin = float(np.abs(O.srgb_to_linear(out) - O.srgb_to_linear(ref)).max())
        srgb = float(np.abs(out - ref).max())
        assert lin < plans["tol"] and srgb < plans["tol"], (m, shape, srgb, lin)
    # the pair MMA computes the same products in the same order: plans agree to fp32 rounding (usually exactly)
    ptol = 2e-5 if plans["flags"] else 1e-4   # (fp8 correction operands: an fp32-level difference can flip a rounding)
    for m in (4, 8, 12):
        assert np.abs(outs[m] - outs[0]).max() < ptol
    assert np.array_equal(outs["nopairs-flag"], outs[0])
    assert np.abs(outs["split-4"] - outs["split-0"]).max() < ptol


def test_pair_plan_partition_independence(plans):
    """Batches and 7-row-halo bands reproduce the single frame bit for bit on pairs too."""
    for key in (12, "split-4"):
        _partition_independence(plans[key])


def _partition_independence(u):
    xs = np.stack([noise_image(40, 140, 50 + i) for i in range(3)])
    outs = u.forward(xs)
    for i in range(3):
        assert np.array_equal(outs[i], u.forward(xs[i]))
    x = noise_image(90, 150, 4321)
    full = u.forward(x)
    parts = []
    for b in range(3):
        y0, y1 = b * 30, (b + 1) * 30
        top, bot = min(7, y0), min(7, 90 - y1)
        parts.append(u.forward_band(x[y0 - top:y1 + bot], top, bot))
    assert np.array_equal(np.concatenate(parts, axis=0), full)


def test_pair_plan_golden_and_u8(plans, weights):
    import rusty_sr_b200 as R
    u = R.Upscaler(weights["anime"], flags=plans["flags"], pair_mask=12)
    lr, hr = load_png("cartoon_lr"), load_png("cartoon_rsa")
    img = u.forward_u8(lr)
    u.close()
    d = np.abs(img.astype(int) - hr.astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= (0.999 if plans["flags"] else 0.995) and (img[..., 3] == 255).all()
