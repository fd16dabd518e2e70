"""ctypes front-end of the CPU oracle (oracle/sr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under rusty_sr_b200/ may import this module.

The shared object is built with -march=native, so it is rebuilt whenever the host CPU's
flag set differs from the one it was built on (the build box and the GPU box may differ).
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsr_oracle.so")
_STAMP = os.path.join(_HERE, ".libsr_oracle.stamp")
_lock = threading.Lock()
_lib = None


def _cpu_stamp() -> str:
    flags = ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    src = b""
    for name in ("sr_oracle.c", "sr_oracle_body.inc", "Makefile"):
        with open(os.path.join(_HERE, name), "rb") as fh:
            src += fh.read()
    return hashlib.sha256(flags.encode() + src).hexdigest()


def build(force: bool = False) -> str:
    """Compile oracle/libsr_oracle.so (gcc -O3 -march=native -fopenmp).  Returns its path."""
    stamp = _cpu_stamp()
    have = None
    if os.path.exists(_STAMP) and os.path.exists(_SO):
        with open(_STAMP) as fh:
            have = fh.read().strip()
    if force or have != stamp:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-B", "CC=gcc"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        with open(_STAMP, "w") as fh:
            fh.write(stamp)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            build()
            L = ctypes.CDLL(_SO)
            fp = ctypes.POINTER(ctypes.c_float)
            dp = ctypes.POINTER(ctypes.c_double)
            L.sr_oracle_num_params.restype = ctypes.c_size_t
            L.sr_oracle_num_params.argtypes = [ctypes.c_int]
            L.sr_oracle_max_threads.restype = ctypes.c_int
            L.sr_oracle_set_threads.argtypes = [ctypes.c_int]
            L.sr_oracle_decode_rsr.restype = ctypes.c_long
            L.sr_oracle_decode_rsr.argtypes = [ctypes.c_char_p, ctypes.c_size_t, fp, ctypes.c_size_t]
            for suf, op in (("f32", fp), ("f64", dp)):
                fn = getattr(L, f"sr_oracle_forward_taps_{suf}")
                fn.restype = ctypes.c_int
                fn.argtypes = [fp, ctypes.c_size_t, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int,
                               op, ctypes.c_void_p]
            for name in ("sr_oracle_bilinear_f64", "sr_oracle_downsample_f64"):
                getattr(L, name).argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp]
                getattr(L, name).restype = None
            L.sr_oracle_conv2d_f64.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int, dp]
            L.sr_oracle_conv2d_f64.restype = ctypes.c_int
            L.sr_oracle_img_to_data.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, fp]
            L.sr_oracle_data_to_img.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
            _lib = L
    return _lib


def num_params(factor: int = 3) -> int:
    return int(lib().sr_oracle_num_params(factor))


def max_threads() -> int:
    return int(lib().sr_oracle_max_threads())


def set_threads(n: int) -> None:
    lib().sr_oracle_set_threads(int(n))


def decode_rsr(data: bytes) -> np.ndarray:
    """bytevec `<Vec<f32>>::decode::<u32>` (reference main.rs:138,146,149,152)."""
    n = lib().sr_oracle_decode_rsr(data, len(data), None, 0)
    if n < 0:
        raise ValueError("ByteVec conversion failed")
    out = np.empty(n, dtype=np.float32)
    rc = lib().sr_oracle_decode_rsr(data, len(data),
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n)
    assert rc == n
    return out


_TAP_NAMES = ("f_activ", "l1_activ", "l2_activ", "l3_activ", "l2_part_f", "l3_part_f", "expand")


def forward(params: np.ndarray, image_hwc3: np.ndarray, factor: int = 3, dtype=np.float32,
            taps: bool = False):
    """graph.forward for sr_net(factor, None) (reference main.rs:171).

    image_hwc3: float32 [H][W][3] in 0..1.  Returns out [fH][fW][3] of `dtype`
    (np.float32 = the precision the reference computes in, np.float64 = ground truth);
    with taps=True returns (out, dict of intermediate nodes).
    """
    params = np.ascontiguousarray(params, dtype=np.float32)
    x = np.ascontiguousarray(image_hwc3, dtype=np.float32)
    assert x.ndim == 3 and x.shape[2] == 3
    H, W = x.shape[:2]
    suf = {np.float32: "f32", np.float64: "f64"}[np.dtype(dtype).type]
    ct = ctypes.c_float if suf == "f32" else ctypes.c_double
    out = np.empty((H * factor, W * factor, 3), dtype=dtype)
    tap_arrays = {}
    tap_ptr = None
    if taps:
        ce = 3 * factor * factor
        ptrs = (ctypes.c_void_p * len(_TAP_NAMES))()
        for i, name in enumerate(_TAP_NAMES):
            c = ce if name == "expand" else 32
            tap_arrays[name] = np.empty((H, W, c), dtype=dtype)
            ptrs[i] = tap_arrays[name].ctypes.data
        tap_ptr = ctypes.cast(ptrs, ctypes.c_void_p)
    fp = ctypes.POINTER(ctypes.c_float)
    fn = getattr(lib(), f"sr_oracle_forward_taps_{suf}")
    rc = fn(params.ctypes.data_as(fp), params.size, factor, x.ctypes.data_as(fp), W, H,
            out.ctypes.data_as(ctypes.POINTER(ct)), tap_ptr)
    if rc == -1:
        raise ValueError("Parameters selected do not have the size required by the neural net")
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return (out, tap_arrays) if taps else out


def bilinear(image_hwc3: np.ndarray, factor: int = 3) -> np.ndarray:
    """bilinear_net(factor) (reference network.rs:111-123), fp64.  Unpinned semantics, see sr_oracle.c."""
    x = np.ascontiguousarray(image_hwc3, dtype=np.float32)
    H, W = x.shape[:2]
    out = np.empty((H * factor, W * factor, 3), np.float64)
    lib().sr_oracle_bilinear_f64(x.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), W, H, factor,
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def downsample(image_hwc3: np.ndarray, factor: int = 3) -> np.ndarray:
    """downsample_net(factor) (reference network.rs:125-138), fp64.  Unpinned semantics, see sr_oracle.c."""
    x = np.ascontiguousarray(image_hwc3, dtype=np.float32)
    H, W = x.shape[:2]
    out = np.empty((-(-H // factor), -(-W // factor), 3), np.float64)
    lib().sr_oracle_downsample_f64(x.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), W, H, factor,
                                   out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def conv2d(x: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """Zero-padded KxK cross-correlation, x f32 [H][W][Ci], weights f32 [Co][K][K][Ci] -> f64 [H][W][Co]
    (alumina Convolution with Padding::Same; reference network.rs:33,60-72)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(weights, dtype=np.float32)
    H, W, ci = x.shape
    co, k = w.shape[0], w.shape[1]
    out = np.empty((H, W, co), np.float64)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = lib().sr_oracle_conv2d_f64(x.ctypes.data_as(fp), W, H, ci, w.ctypes.data_as(fp), co, k,
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    if rc:
        raise MemoryError("oracle conv2d failed")
    return out


def img_to_data(img_u8: np.ndarray) -> np.ndarray:
    """alumina img_to_data (reference main.rs:170): u8 RGB(A) -> f32 [H][W][3] = u8/255."""
    img = np.ascontiguousarray(img_u8, dtype=np.uint8)
    H, W, C = img.shape
    out = np.empty((H, W, 3), dtype=np.float32)
    lib().sr_oracle_img_to_data(img.ctypes.data_as(ctypes.c_char_p), W, H, C,
                                out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def data_to_img(data: np.ndarray) -> np.ndarray:
    """alumina data_to_img + .to_rgba() (reference main.rs:175): clamp, *255, round, alpha=255."""
    d = np.ascontiguousarray(data, dtype=np.float32)
    H, W, _ = d.shape
    out = np.empty((H, W, 4), dtype=np.uint8)
    lib().sr_oracle_data_to_img(d.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), W, H,
                                out.ctypes.data_as(ctypes.c_char_p))
    return out


def srgb_to_linear(v: np.ndarray) -> np.ndarray:
    """sRGB EOTF used for the 'linear light' parity metric (BASELINE.json north_star)."""
    v = np.clip(np.asarray(v, dtype=np.float64), 0.0, 1.0)
    return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)
