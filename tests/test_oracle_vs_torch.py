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
