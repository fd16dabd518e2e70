"""Second, independent restatement of the path — plain PyTorch ops on the CPU in fp64 (conv2d, pixel_shuffle, bilinear
interpolate) written from SURVEY.md Appendix A, not from oracle/sr_oracle.c — cross-checked against the C oracle.  Two
implementations that share no code agreeing to 1e-11 pins the oracle's arithmetic beyond the two golden images
(reference graph: src/network.rs:27-72; op semantics: SURVEY.md A.2 / A.3)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_png, noise_image
from oracle import oracle as O


def torch_forward(params: np.ndarray, x_hwc3: np.ndarray, factor: int = 3) -> np.ndarray:
    import rusty_sr_b200 as R   # parameter layout only (names / offsets / shapes of network.rs's op order)
    p = {}
    for s in R.param_layout():
        n = int(np.prod(s.shape))
        p[s.name] = torch.from_numpy(params[s.offset:s.offset + n].astype(np.float64).reshape(s.shape))
    conv = lambda t, name: F.conv2d(t, p[name].permute(0, 3, 1, 2), padding=p[name].shape[1] // 2)   # [Cout][Kh][Kw][Cin], no flip, zero pad
    belu = lambda t, name: p[name].view(1, -1, 1, 1) * t + torch.sqrt(t * t + 1.0) - 1.0              # BeLU(x; p) = p x + sqrt(x^2 + 1) - 1
    bias = lambda name: p[name].view(1, -1, 1, 1)
    x = torch.from_numpy(x_hwc3.astype(np.float64)).permute(2, 0, 1)[None]
    f = belu(conv(x, "conv0") + bias("f_bias"), "f_activ")
    l1 = belu(conv(f, "conv1") + bias("l1_bias"), "l1_activ")
    l2 = belu(conv(f, "conv2") + conv(l1, "conv5") + bias("l2_bias"), "l2_activ")
    l3 = belu(conv(f, "conv3") + conv(l1, "conv6") + conv(l2, "conv8") + bias("l3_bias"), "l3_activ")
    e = conv(l1, "conv7") + conv(l2, "conv9") + conv(l3, "conv10") + bias("expand_bias")
    # Expand: out[f*h+dy, f*w+dx, c] = e[h, w, (dy*f + dx)*3 + c]  -> pixel_shuffle wants channel index c*f*f + dy*f + dx
    n, _, H, W = e.shape
    e = e.view(n, factor * factor, 3, H, W).permute(0, 2, 1, 3, 4).reshape(n, 3 * factor * factor, H, W)
    out = F.pixel_shuffle(e, factor) + F.interpolate(x, scale_factor=factor, mode="bilinear", align_corners=False)
    return out[0].permute(1, 2, 0).numpy()


@pytest.mark.parametrize("wname", ["imagenet", "imagenetlinear", "anime"])
@pytest.mark.parametrize("case", ["cartoon", "noise", "tiny"])
def test_c_oracle_equals_torch_restatement(weights, wname, case):
    torch.set_num_threads(4)
    x = {"cartoon": lambda: O.img_to_data(load_png("cartoon_lr"))[:48, :64], "noise": lambda: noise_image(37, 53, 99),
         "tiny": lambda: noise_image(1, 2, 5)}[case]()
    ref = torch_forward(weights[wname], x)
    got = O.forward(weights[wname], x, 3, np.float64)
    assert got.shape == ref.shape == (x.shape[0] * 3, x.shape[1] * 3, 3)
    assert np.abs(got - ref).max() < 1e-11
    assert np.abs(O.forward(weights[wname], x, 3, np.float32).astype(np.float64) - ref).max() < 5e-5


def _srgb_to_linear(t):
    return torch.where(t <= 0.04045, t / 12.92, ((t + 0.055) / 1.055) ** 2.4)


def _linear_to_srgb(t):
    return torch.where(t <= 0.0031308, t * 12.92, 1.055 * t ** (1.0 / 2.4) - 0.055)


@pytest.mark.parametrize("shape", [(12, 15), (1, 1), (7, 5), (40, 31)])
def test_parameter_free_modes_equal_torch_restatement(shape):
    """bilinear_net / downsample_net (network.rs:111-138): SrgbToLinear -> LinearInterp | Pooling -> LinearToSrgb.  (Internal
    consistency of two restatements; the reference commits no output of these modes, so they stay "parity unpinned".)"""
    x = noise_image(shape[0], shape[1], 17)
    t = torch.from_numpy(x.astype(np.float64)).permute(2, 0, 1)[None]
    up = _linear_to_srgb(F.interpolate(_srgb_to_linear(t), scale_factor=3, mode="bilinear", align_corners=False))
    assert np.abs(O.bilinear(x, 3) - up[0].permute(1, 2, 0).numpy()).max() < 1e-12
    dn = _linear_to_srgb(F.avg_pool2d(_srgb_to_linear(t), 3, ceil_mode=True, count_include_pad=False))
    assert np.abs(O.downsample(x, 3) - dn[0].permute(1, 2, 0).numpy()).max() < 1e-12
