"""The compiled host program (rusty_sr_b200/host/rusty_sr.cpp -> rusty_sr_b200/bin/rusty_sr) mirrors the
reference CLI (src/main.rs:33-178): flags, conflicts, messages, .rsr handling, PNG in / RGBA PNG out."""
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

from conftest import GOLDEN, RES, ROOT, load_png

BIN = os.path.join(ROOT, "rusty_sr_b200", "bin", "rusty_sr")


def run(*args):
    return subprocess.run([BIN, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(BIN):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.run(["make", "-C", os.path.join(ROOT, "rusty_sr_b200", "csrc"), "host"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def test_flag_surface_matches_reference_clap_definitions():
    assert run().returncode == 2                                       # INPUT_FILE / OUTPUT_FILE required (main.rs:39-48)
    assert run("a.png", "b.png", "-p", "nope").returncode == 2         # possible_values (main.rs:55)
    assert run("a.png", "b.png", "-p", "anime", "-c", "x.rsr").returncode == 2   # conflicts_with (main.rs:59)
    assert run("a.png", "b.png", "-d", "-p", "anime").returncode == 2  # conflicts_with (main.rs:69-70)
    r = run("a.png", "b.png", "-c", "/nonexistent.rsr")
    assert r.returncode == 101 and "Error opening parameter file" in r.stderr      # main.rs:134


def test_custom_parameter_file_errors(tmp_path):
    bad = tmp_path / "bad.rsr"
    bad.write_bytes(open(os.path.join(RES, "anime.rsr"), "rb").read()[:-8])
    r = run(os.path.join(GOLDEN, "cartoon_lr.png"), str(tmp_path / "o.png"), "-c", str(bad))
    assert r.returncode == 101 and "ByteVec conversion failed" in r.stderr         # main.rs:138
    short = tmp_path / "short.rsr"
    import rusty_sr_b200 as R
    short.write_bytes(R.encode_rsr(np.zeros(10, np.float32)))
    r = run(os.path.join(GOLDEN, "cartoon_lr.png"), str(tmp_path / "o.png"), "-c", str(short))
    assert r.returncode == 101 and "do not have the size required by the neural net" in r.stderr  # main.rs:162


def test_parameter_free_modes_print_the_reference_messages(tmp_path):
    import torch
    r = run(os.path.join(GOLDEN, "cartoon_lr.png"), str(tmp_path / "o.png"), "-p", "bilinear")
    assert "Upscaling using bilinear interpolation..." in r.stdout                # main.rs:154
    assert r.returncode == (0 if torch.cuda.is_available() else 101)
    r = run(os.path.join(GOLDEN, "cartoon_lr.png"), str(tmp_path / "o2.png"), "-d")
    assert "Downsampling using average pooling of linear RGB values..." in r.stdout  # main.rs:140
    assert r.returncode == (0 if torch.cuda.is_available() else 101)


@pytest.mark.gpu
def test_cli_bilinear_and_downsample(tmp_path):
    from oracle import oracle as O
    lr = load_png("cartoon_lr")
    for flag, ofn in ((["-p", "bilinear"], O.bilinear), (["-d"], O.downsample)):
        out = tmp_path / "o.png"
        assert run(os.path.join(GOLDEN, "cartoon_lr.png"), str(out), *flag).returncode == 0
        got = np.asarray(Image.open(out))
        want = O.data_to_img(ofn(O.img_to_data(lr), 3).astype(np.float32))
        assert got.shape == want.shape and np.abs(got.astype(int) - want.astype(int)).max() <= 1


@pytest.mark.parametrize("name", ["cartoon_lr", "butterfly_lr", "cartoon_rsa"])
def test_png_codec_round_trip(tmp_path, name):
    out = tmp_path / "rt.png"
    r = run("--png-roundtrip", os.path.join(GOLDEN, name + ".png"), str(out))
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.asarray(Image.open(out).convert("RGBA")), load_png(name))


def test_png_reader_handles_rgb_grey_and_palette(tmp_path):
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, size=(13, 17, 3), dtype=np.uint8)
    for mode, arr in (("RGB", rgb), ("L", rgb[..., 0]), ("P", None), ("LA", None)):
        src = tmp_path / f"in_{mode}.png"
        im = Image.fromarray(rgb, "RGB")
        if mode == "L":
            im = Image.fromarray(arr, "L")
        elif mode == "P":
            im = im.convert("P", palette=Image.ADAPTIVE, colors=16)
        elif mode == "LA":
            im = im.convert("LA")
        im.save(src)
        out = tmp_path / f"out_{mode}.png"
        assert run("--png-roundtrip", str(src), str(out)).returncode == 0
        assert np.array_equal(np.asarray(Image.open(out)), np.asarray(Image.open(src).convert("RGBA")))


def test_no_gpu_is_a_loud_failure(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run(os.path.join(GOLDEN, "cartoon_lr.png"), str(tmp_path / "o.png"), "-p", "anime")
    assert r.returncode == 101 and "no CUDA device" in r.stderr
    assert "Upscaling using anime neural net parameters..." in r.stdout           # main.rs:151


@pytest.mark.gpu
def test_cli_reproduces_the_reference_golden_image(tmp_path):
    """rusty_sr docs/cartoon_lr.png out.png -p anime  ==  docs/cartoon_rsa.png (README.md:24), bit level."""
    out = tmp_path / "cartoon.png"
    r = run(os.path.join(GOLDEN, "cartoon_lr.png"), str(out), "-p", "anime")
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "Upscaling using anime neural net parameters... Writing file... Done"   # main.rs:151,173,177
    got, want = np.asarray(Image.open(out)), load_png("cartoon_rsa")
    d = np.abs(got.astype(int) - want.astype(int))
    assert got.shape == want.shape and d.max() <= 1 and (d == 0).mean() >= 0.999
    out2 = tmp_path / "butterfly.png"
    assert run(os.path.join(GOLDEN, "butterfly_lr.png"), str(out2)).returncode == 0   # default parameters: imagenet (main.rs:144)
    got, want = np.asarray(Image.open(out2)), load_png("butterfly_rs")
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 5


@pytest.mark.gpu
def test_python_cli_matches_compiled_cli(tmp_path):
    """python -m rusty_sr_b200.cli has the same flag surface and produces the same image as bin/rusty_sr."""
    import sys
    a, b = tmp_path / "a.png", tmp_path / "b.png"
    src = os.path.join(GOLDEN, "cartoon_lr.png")
    assert run(src, str(a), "-p", "imagenetlinear").returncode == 0
    r = subprocess.run([sys.executable, "-m", "rusty_sr_b200.cli", src, str(b), "-p", "imagenetlinear"], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Upscaling using linear loss imagenet neural net parameters..." in r.stdout      # main.rs:148
    assert np.array_equal(np.asarray(Image.open(a)), np.asarray(Image.open(b)))
