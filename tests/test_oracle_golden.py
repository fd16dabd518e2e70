"""Pins the CPU oracle against everything the reference commits that fixes numeric results:
the two README image pairs (docs/*_lr.png -> docs/*_rs*.png), the parameter-count assert
(main.rs:162), the .rsr blob format, and the fp64 known-answer values of SURVEY.md App. B.4.
The reference has no tests of its own."""
import hashlib
import os

import numpy as np
import pytest

from conftest import RES, load_png, synthetic_kat_image
from oracle import oracle as O


def test_num_params_matches_reference_assert():
    # main.rs:162 assert_eq!(params.len(), graph.num_params()) with the bundled 130459-float blobs
    assert O.num_params(3) == 130459
    assert O.num_params(2) == 130459 - 3 * (27 - 12) * (9 * 32) - (27 - 12)
    assert O.num_params(4) == 130459 + 3 * (48 - 27) * (9 * 32) + (48 - 27)


@pytest.mark.parametrize("name,sha", [
    ("anime", "aaf006212995ede02eed1b8113da9f690ccb246ed225ccedbf2a8df57bb5b2b0"),
    ("imagenet", "5781fba8676d626b26e9e3f787bcdcb5b3d77ffedf4fb5f5102337720c551e3c"),
    ("imagenetlinear", "4e37c3e1e4f0cf7646550a7a95f44086b9587feac20bc8f54587f5d614649e59"),
])
def test_rsr_blobs(name, sha):
    data = open(os.path.join(RES, name + ".rsr"), "rb").read()
    assert hashlib.sha256(data).hexdigest() == sha
    assert len(data) == 4 + 8 * 130459
    p = O.decode_rsr(data)
    assert p.shape == (130459,) and np.isfinite(p).all()


def test_rsr_decode_rejects_malformed():
    data = open(os.path.join(RES, "anime.rsr"), "rb").read()
    with pytest.raises(ValueError):
        O.decode_rsr(data[:-4])
    with pytest.raises(ValueError):
        O.decode_rsr(b"\x01\x00")
    bad = bytearray(data)
    bad[4] = 8  # an element size that is not 4
    with pytest.raises(ValueError):
        O.decode_rsr(bytes(bad))


def test_param_count_mismatch_is_an_error(weights):
    with pytest.raises(ValueError):
        O.forward(weights["anime"][:-1], np.zeros((4, 4, 3), np.float32))
    with pytest.raises(ValueError):
        O.forward(weights["anime"], np.zeros((4, 4, 3), np.float32), factor=4)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_golden_cartoon_bit_level(weights, dtype):
    """docs/cartoon_lr.png + anime.rsr -> docs/cartoon_rsa.png (README.md:24): bit-level."""
    lr, hr = load_png("cartoon_lr"), load_png("cartoon_rsa")
    assert lr.shape == (120, 84, 4) and hr.shape == (360, 252, 4)
    out = O.forward(weights["anime"], O.img_to_data(lr), 3, dtype)
    img = O.data_to_img(out.astype(np.float32))
    d = np.abs(img.astype(int) - hr.astype(int))
    assert d.max() <= 1
    assert (d == 0).mean() >= 0.9999
    assert (img[..., 3] == 255).all()
    # SURVEY.md B.4 known answers (fp64 probe)
    tol = 1e-6 if dtype == np.float64 else 2e-5
    np.testing.assert_allclose(out[0, 0], [0.5964677691, 0.6237049025, 0.6240415982], atol=tol)
    np.testing.assert_allclose(out[100, 100], [0.6324843026, 0.6098251676, 0.6122859687], atol=tol)
    np.testing.assert_allclose(out[-1, -1], [0.2945901404, 0.2129077157, 0.0799281953], atol=tol)
    assert abs(out.min() - (-0.0356005436)) < tol and abs(out.max() - 1.1244185727) < tol
    assert abs(out.sum(dtype=np.float64) - 148627.07815674) < 0.05


def test_golden_butterfly_55db(weights):
    """docs/butterfly_lr.png + imagenet.rsr -> docs/butterfly_rs.png (README.md:27):
    made with slightly older weights, so >= 55 dB and <= 5 LSB is the gate."""
    lr, hr = load_png("butterfly_lr"), load_png("butterfly_rs")
    out = O.forward(weights["imagenet"], O.img_to_data(lr), 3, np.float32)
    img = O.data_to_img(out)
    d = np.abs(img.astype(int) - hr.astype(int))
    mse = np.mean((img[..., :3].astype(float) - hr[..., :3]) ** 2)
    assert 10 * np.log10(255 ** 2 / mse) >= 55.0
    assert d.max() <= 5
    np.testing.assert_allclose(out[0, 0], [0.4260106917, 0.3691584805, 0.1800157537], atol=2e-5)
    np.testing.assert_allclose(out[-1, -1], [0.5188002834, 0.3084579738, 0.2366696892], atol=2e-5)
    assert abs(out.sum(dtype=np.float64) - 480272.83269361) < 0.1


@pytest.mark.parametrize("name,total,lo,hi,s00,s2431", [
    ("imagenet", 3992.76461685, -0.31857803, 1.40157759,
     [-0.0265731391, -0.0054821758, 0.0289644812], [0.5922541513, 0.6175470251, 0.6417088860]),
    ("imagenetlinear", 4224.45130896, -0.40361350, 1.18794873,
     [-0.0291555039, -0.0041451522, 0.0236015430], [0.5814859209, 0.6188186111, 0.6513154810]),
    ("anime", 3994.37217562, -0.34765356, 1.44840150,
     [-0.0259854044, -0.0040771457, 0.0316297421], [0.6039646963, 0.6256366248, 0.6451362323]),
])
def test_known_answers_synthetic(weights, name, total, lo, hi, s00, s2431):
    out = O.forward(weights[name], synthetic_kat_image(), 3, np.float64)
    assert out.shape == (48, 60, 3)
    assert abs(out.sum() - total) < 1e-3
    assert abs(out.min() - lo) < 1e-6 and abs(out.max() - hi) < 1e-6
    np.testing.assert_allclose(out[0, 0], s00, atol=1e-6)
    np.testing.assert_allclose(out[24, 31], s2431, atol=1e-6)


def test_fp32_oracle_close_to_fp64(weights):
    x = synthetic_kat_image(33, 21)
    a = O.forward(weights["imagenet"], x, 3, np.float32)
    b = O.forward(weights["imagenet"], x, 3, np.float64)
    assert np.abs(a - b).max() < 5e-5


def test_null_image_residual(weights):
    """all-zero input -> interior residual |r| <= 6.5e-4 (a wrong activation
    gives 0.05..20)."""
    out = O.forward(weights["imagenet"], np.zeros((24, 24, 3), np.float32), 3, np.float64)
    # interior = more than the +-7 px receptive field away from the zero-padded border
    assert np.abs(out[24:48, 24:48]).max() < 7e-4
    np.testing.assert_allclose(out[34, 34:37, 0], [0.00038281, 0.00036596, 0.00028766], atol=2e-6)


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (7, 5), (2, 130)])
def test_odd_sizes_and_taps(weights, shape):
    H, W = shape
    rng = np.random.default_rng(7)
    x = rng.random((H, W, 3), dtype=np.float32)
    out, taps = O.forward(weights["anime"], x, 3, np.float64, taps=True)
    assert out.shape == (3 * H, 3 * W, 3) and np.isfinite(out).all()
    assert taps["f_activ"].shape == (H, W, 32) and taps["expand"].shape == (H, W, 27)


def test_u8_mapping_round_trip():
    img = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(4, axis=2)
    x = O.img_to_data(img)
    assert x.shape == (16, 16, 3) and x.max() == 1.0
    back = O.data_to_img(x)
    assert (back[..., :3] == img[..., :3]).all() and (back[..., 3] == 255).all()
    assert (O.data_to_img(np.full((1, 1, 3), 1.7, np.float32))[0, 0, :3] == 255).all()
    assert (O.data_to_img(np.full((1, 1, 3), -0.2, np.float32))[0, 0, :3] == 0).all()
