"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/rsr_b200.h declares, and maps the reference's failure modes (main.rs:134-138,162) to
error codes without needing a GPU.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import RES, ROOT
import rusty_sr_b200 as R
from rusty_sr_b200 import _lib


def header_functions():
    text = open(os.path.join(ROOT, "include", "rsr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    declared = header_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/rsr_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared


def test_num_params_matches_reference_assert():
    L = _lib.load()
    assert L.rsr_num_params(3) == 130459  # main.rs:162 with the bundled blobs
    assert R.num_params(3) == 130459
    assert L.rsr_num_params(0) == 0 and L.rsr_num_params(5) == 0
    spans = R.param_layout(3)
    assert [s.name for s in spans][:4] == ["conv0", "f_bias", "f_activ", "expand_bias"]
    off = {s.name: s.offset for s in spans}
    # SURVEY.md A.2 offsets
    assert off["conv1"] == 2683 and off["conv5"] == 79483 and off["conv10"] == 122683
    assert off["l1_bias"] == 2491 and off["l3_activ"] == 2651


def test_param_count_mismatch_is_reported_like_the_reference():
    L = _lib.load()
    p = np.zeros(100, np.float32)
    h = L.rsr_create(p.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), p.size, 3, 0, 0)
    assert not h
    assert L.rsr_last_error_code() == _lib.RSR_E_PARAM_COUNT
    assert b"do not have the size required by the neural net" in L.rsr_last_error()
    with pytest.raises(ValueError):
        R.Upscaler(params=p)


def test_bad_arguments():
    L = _lib.load()
    assert not L.rsr_create(None, 0, 3, 0, 0)
    assert L.rsr_last_error_code() == _lib.RSR_E_BAD_ARG
    p = np.zeros(R.num_params(4), np.float32)
    pp = p.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    assert not L.rsr_create(pp, p.size, 5, 0, 0)                       # factors 1..4 only
    assert L.rsr_last_error_code() == _lib.RSR_E_BAD_ARG
    assert not L.rsr_create(pp, p.size, 4, 0, _lib.RSR_FLAG_FUSED_MID)  # the fused middle launch holds 32 expand channels
    assert L.rsr_last_error_code() == _lib.RSR_E_UNSUPPORTED
    assert [L.rsr_num_params(f) for f in (1, 2, 3, 4)] == [R.num_params(f) for f in (1, 2, 3, 4)]
    assert L.rsr_upscale_f32(None, None, 1, 1, None) == _lib.RSR_E_BAD_ARG
    assert L.rsr_synchronize(None) == _lib.RSR_E_BAD_ARG


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(R.RsrError) as ei:
        R.Upscaler(parameters="anime")
    assert ei.value.code == _lib.RSR_E_CUDA


def test_rsr_codec_round_trip_and_errors():
    data = open(os.path.join(RES, "imagenetlinear.rsr"), "rb").read()
    p = R.decode_rsr(data)
    assert p.dtype == np.float32 and p.size == 130459
    assert R.encode_rsr(p) == data
    for bad in (b"", b"\x05\x00\x00\x00", data[:-1], data + b"\x00"):
        with pytest.raises(ValueError, match="ByteVec conversion failed"):
            R.decode_rsr(bad)
    # agrees with the oracle's decoder
    from oracle import oracle as O
    assert np.array_equal(O.decode_rsr(data), p)
    # and with the codec of the C ABI (rsr_rsr_encode / rsr_rsr_decode: what the compiled host uses; main.rs:138,213)
    L = _lib.load()
    buf = np.frombuffer(data, np.uint8)
    assert L.rsr_rsr_decode(buf.ctypes.data, buf.size, None, 0) == p.size
    q = np.zeros(p.size, np.float32)
    assert L.rsr_rsr_decode(buf.ctypes.data, buf.size, q.ctypes.data, q.size) == p.size and np.array_equal(q, p)
    out = np.zeros(L.rsr_rsr_encoded_size(p.size), np.uint8)
    assert out.size == len(data) and L.rsr_rsr_encode(p.ctypes.data, p.size, out.ctypes.data, out.size) == len(data)
    assert out.tobytes() == data
    assert L.rsr_rsr_encode(p.ctypes.data, p.size, out.ctypes.data, out.size - 1) == _lib.RSR_E_BAD_ARG
    for bad in (data[:-1], data + b"\x00", b"\x05\x00\x00\x00"):
        b = np.frombuffer(bad, np.uint8)
        assert L.rsr_rsr_decode(b.ctypes.data, b.size, None, 0) == _lib.RSR_E_BAD_ARG


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "rusty_sr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower().replace("no cpu or pytorch fallback", ""), f
