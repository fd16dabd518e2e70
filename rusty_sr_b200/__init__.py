"""rusty_sr_b200 — B200-native forward path of millardjn/rusty_sr's super-resolution net.

Host-side mirror of the reference's `upscale()` (src/main.rs:130-178): pick parameters
(`-p imagenet|imagenetlinear|anime`, `-c FILE.rsr`), decode the .rsr blob, map the image to
the tensor `graph.forward` takes, run the forward pass, map back.  The forward pass itself is
the C-ABI library `lib/librsr_b200.so` (hand-written sm_100a kernels; include/rsr_b200.h).
There is no CPU fallback: importing works anywhere, running needs a B200.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import RsrError, RSR_FLAG_FUSED_MID, RSR_FLAG_SPLIT_MID, RSR_FLAG_NO_PAIRS, RSR_FLAG_BF16X3
from .network import FACTOR, num_params, param_layout
from .rsr_format import decode_rsr, encode_rsr

__all__ = ["Upscaler", "MultiUpscaler", "upscale", "bilinear", "downsample", "conv2d_3x3", "load_builtin_params", "decode_rsr", "encode_rsr", "RsrError",
           "num_params", "param_layout", "FACTOR", "BUILTIN_PARAMETERS", "RSR_FLAG_FUSED_MID", "RSR_FLAG_SPLIT_MID", "RSR_FLAG_NO_PAIRS", "RSR_FLAG_BF16X3"]

_RES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "res")

# main.rs:26-28 include_bytes!("res/*.rsr"); `-p` possible values main.rs:55
BUILTIN_PARAMETERS = ("imagenet", "imagenetlinear", "anime")


def load_builtin_params(name: str = "imagenet") -> np.ndarray:
    """The parameter vectors the reference embeds (main.rs:143-152; default imagenet :144)."""
    if name not in BUILTIN_PARAMETERS:
        raise ValueError(f"unknown built-in parameters {name!r}; choose from {BUILTIN_PARAMETERS}")
    with open(os.path.join(_RES, name + ".rsr"), "rb") as fh:
        return decode_rsr(fh.read())


def _as_ptr(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


class Upscaler:
    """One network instance bound to one GPU (the reference's `(params, graph)` pair,
    main.rs:133-158).  Not thread-safe, like the reference's `&mut Graph`."""

    def __init__(self, params: Optional[np.ndarray] = None, parameters: str = "imagenet",
                 factor: int = FACTOR, device: int = 0, flags: int = 0, pair_mask: Optional[int] = None):
        """`flags`: RSR_FLAG_FUSED_MID / RSR_FLAG_SPLIT_MID select the launch plan of the middle levels, RSR_FLAG_BF16X3
        the three-product operand split (include/rsr_b200.h); 0 = library default.  `pair_mask` (tests only): which
        launches run on CTA pairs (rsr_debug_create)."""
        if params is None:
            params = load_builtin_params(parameters)
        params = np.ascontiguousarray(params, dtype=np.float32).ravel()
        self._L = _lib.load()
        self.factor = factor
        self.device = device
        self.n_params = int(params.size)
        pp = params.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        if pair_mask is None:
            h = self._L.rsr_create(pp, params.size, factor, device, int(flags))
        else:
            h = self._L.rsr_debug_create(pp, params.size, factor, device, int(flags), int(pair_mask))
        if not h:
            code = self._L.rsr_last_error_code()
            msg = self._L.rsr_last_error().decode("utf-8", "replace")
            if code == _lib.RSR_E_PARAM_COUNT:
                # the reference's assert_eq! message (main.rs:162)
                raise ValueError(msg)
            raise RsrError(code, msg)
        self._h = ctypes.c_void_p(h)

    # -- lifetime -------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.rsr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- graph.forward --------------------------------------------------------------------
    def forward(self, x_hwc3: np.ndarray) -> np.ndarray:
        """graph.forward(1, vec![input], &params).remove(0) (main.rs:171).
        x: f32 [H][W][3] (or [N][H][W][3]) sRGB 0..1 -> f32 [fH][fW][3], not clamped."""
        x = np.ascontiguousarray(x_hwc3, dtype=np.float32)
        if x.ndim == 3:
            if x.shape[2] != 3:
                raise ValueError("expected [H][W][3]")
            H, W = x.shape[:2]
            out = np.empty((H * self.factor, W * self.factor, 3), np.float32)
            _lib.check(self._L.rsr_upscale_f32(self._h, _as_ptr(x), W, H, _as_ptr(out)))
            return out
        if x.ndim == 4 and x.shape[3] == 3:
            N, H, W = x.shape[:3]
            out = np.empty((N, H * self.factor, W * self.factor, 3), np.float32)
            _lib.check(self._L.rsr_upscale_batch_f32(self._h, _as_ptr(x), N, W, H, _as_ptr(out)))
            return out
        raise ValueError("expected [H][W][3] or [N][H][W][3]")

    def forward_u8(self, img: np.ndarray, rgb: bool = False) -> np.ndarray:
        """img_to_data -> forward -> data_to_img(..).to_rgba() fused on the device
        (main.rs:170-175).  img: u8 [H][W][3|4] or [N][H][W][3|4] -> u8 RGBA [.., fH, fW, 4];
        rgb=True (single images): u8 RGB [fH][fW][3], i.e. data_to_img without the constant alpha."""
        a = np.ascontiguousarray(img, dtype=np.uint8)
        if a.ndim == 3:
            H, W, C = a.shape
            out = np.empty((H * self.factor, W * self.factor, 3 if rgb else 4), np.uint8)
            fn = self._L.rsr_upscale_u8_rgb if rgb else self._L.rsr_upscale_u8
            _lib.check(fn(self._h, _as_ptr(a), W, H, C, _as_ptr(out)))
            return out
        if a.ndim == 4:
            N, H, W, C = a.shape
            out = np.empty((N, H * self.factor, W * self.factor, 4), np.uint8)
            _lib.check(self._L.rsr_upscale_batch_u8(self._h, _as_ptr(a), N, W, H, C, _as_ptr(out)))
            return out
        raise ValueError("expected [H][W][C] or [N][H][W][C]")

    # -- pipelined host calls: copies of neighbouring frames overlap the compute -----------------
    def _check_frame(self, a: np.ndarray, out: np.ndarray, dtype, in_channels, out_channels: int) -> tuple:
        # explicit checks (not asserts): a wrongly sized `out` would let the asynchronous copy write past the buffer
        for name, t in (("input", a), ("out", out)):
            if not isinstance(t, np.ndarray) or t.dtype != dtype or not t.flags.c_contiguous:
                raise ValueError(f"{name} must be a C-contiguous numpy array of dtype {np.dtype(dtype).name}")
        if a.ndim != 3 or a.shape[2] not in in_channels:
            raise ValueError(f"expected an [H][W][{'|'.join(map(str, in_channels))}] frame, got shape {a.shape}")
        H, W, C = a.shape
        want = (H * self.factor, W * self.factor, out_channels)
        if out.shape != want:
            raise ValueError(f"out must have shape {want}, got {out.shape}")
        return H, W, C

    def submit_u8(self, img: np.ndarray, out: np.ndarray) -> int:
        """Enqueue one frame (u8 [H][W][3|4] -> u8 `out`, both C-contiguous and ideally pinned); `out` is RGBA
        [fH][fW][4] (data_to_img(..).to_rgba(), main.rs:175) or RGB [fH][fW][3] (data_to_img alone: the alpha the
        reference adds on the host is constant 255).  Returns a ticket for wait(); buffers must stay alive until then."""
        oc = out.shape[2] if isinstance(out, np.ndarray) and out.ndim == 3 and out.shape[2] in (3, 4) else 4
        H, W, C = self._check_frame(img, out, np.uint8, (3, 4), oc)
        fn = self._L.rsr_submit_u8 if oc == 4 else self._L.rsr_submit_u8_rgb
        t = fn(self._h, _as_ptr(img), W, H, C, _as_ptr(out))
        if t < 0:
            _lib.check(t)
        return t

    def submit_f32(self, x: np.ndarray, out: np.ndarray) -> int:
        H, W, _ = self._check_frame(x, out, np.float32, (3,), 3)
        t = self._L.rsr_submit_f32(self._h, _as_ptr(x), W, H, _as_ptr(out))
        if t < 0:
            _lib.check(t)
        return t

    def wait(self, ticket: int) -> None:
        _lib.check(self._L.rsr_wait(self._h, ticket))

    def wait_all(self) -> None:
        _lib.check(self._L.rsr_wait_all(self._h))

    def forward_band(self, rows: np.ndarray, top_halo: int, bottom_halo: int) -> np.ndarray:
        """One horizontal band of a larger frame (spatial sharding, overlap-recompute)."""
        x = np.ascontiguousarray(rows, dtype=np.float32)
        H, W = x.shape[:2]
        if top_halo < 0 or bottom_halo < 0 or H - top_halo - bottom_halo < 1:
            raise ValueError(f"bad band: rows_in={H} top_halo={top_halo} bottom_halo={bottom_halo}")
        out = np.empty(((H - top_halo - bottom_halo) * self.factor, W * self.factor, 3), np.float32)
        _lib.check(self._L.rsr_upscale_band_f32(self._h, _as_ptr(x), W, H, top_halo, bottom_halo,
                                                _as_ptr(out)))
        return out

    # -- device-resident entry points (raw device pointers as ints) ------------------------
    def forward_device(self, d_in: int, n: int, w: int, h: int, d_out_f32: int = 0,
                       d_out_rgba: int = 0, u8_channels: int = 0) -> None:
        if u8_channels:
            rc = self._L.rsr_upscale_device_u8(self._h, ctypes.c_void_p(d_in), n, w, h, u8_channels,
                                               ctypes.c_void_p(d_out_f32 or None),
                                               ctypes.c_void_p(d_out_rgba or None))
        else:
            rc = self._L.rsr_upscale_device_f32(self._h, ctypes.c_void_p(d_in), n, w, h,
                                                ctypes.c_void_p(d_out_f32 or None),
                                                ctypes.c_void_p(d_out_rgba or None))
        _lib.check(rc)

    def forward_device_u8_rgb(self, d_in: int, n: int, w: int, h: int, in_channels: int, d_out_rgb: int) -> None:
        """Device-resident u8 [n][h][w][in_channels] -> u8 RGB [n][fh][fw][3] (rsr_upscale_device_u8_rgb); not synchronised."""
        _lib.check(self._L.rsr_upscale_device_u8_rgb(self._h, ctypes.c_void_p(d_in), n, w, h, in_channels, ctypes.c_void_p(d_out_rgb)))

    def forward_band_device(self, d_in_rows: int, w: int, rows_in: int, top_halo: int, bottom_halo: int, d_out: int,
                            u8_channels: int = 0, out_channels: int = 0) -> None:
        """Device-resident band (rsr_upscale_band_device_*): u8_channels = 0 -> f32 in / f32 out, else u8 in / u8 out with
        out_channels 3 (RGB) or 4 (RGBA).  Enqueued on the context's stream, not synchronised."""
        if u8_channels:
            rc = self._L.rsr_upscale_band_device_u8(self._h, ctypes.c_void_p(d_in_rows), w, rows_in, top_halo, bottom_halo,
                                                    u8_channels, ctypes.c_void_p(d_out), out_channels or 4)
        else:
            rc = self._L.rsr_upscale_band_device_f32(self._h, ctypes.c_void_p(d_in_rows), w, rows_in, top_halo, bottom_halo,
                                                     ctypes.c_void_p(d_out))
        _lib.check(rc)

    def set_stream(self, cuda_stream: int) -> None:
        _lib.check(self._L.rsr_set_stream(self._h, ctypes.c_void_p(cuda_stream or None)))

    def synchronize(self) -> None:
        _lib.check(self._L.rsr_synchronize(self._h))

    def set_kernel_timing(self, enable: bool = True) -> None:
        """Record CUDA events around every launch so that last_kernel_times_ms() has data (off by default: the events
        serialise the launches, which otherwise overlap through programmatic dependent launch)."""
        _lib.check(self._L.rsr_set_kernel_timing(self._h, int(enable)))

    def last_kernel_times_ms(self) -> dict:
        buf = (ctypes.c_float * 16)()
        n = self._L.rsr_last_kernel_times(self._h, buf, 16)
        if n < 0:
            _lib.check(n)
        return {self._L.rsr_kernel_name(i).decode(): float(buf[i]) for i in range(n)}

    # -- test hooks (tests/ only) ----------------------------------------------------------
    def debug_stop_after(self, launch_index: int) -> None:
        _lib.check(self._L.rsr_debug_stop_after(self._h, launch_index))

    def debug_read_node(self, which: int, n_pixels: int) -> np.ndarray:
        out = np.empty((n_pixels, 32), np.float32)
        _lib.check(self._L.rsr_debug_read_node(self._h, which, _as_ptr(out), n_pixels))
        return out

    def debug_profile(self, enable: bool = True) -> None:
        _lib.check(self._L.rsr_debug_profile(self._h, int(enable)))

    def debug_read_profile(self, launch_index: int) -> np.ndarray:
        out = np.zeros((256, 16), np.int64)
        n = self._L.rsr_debug_read_profile(self._h, launch_index, _as_ptr(out), 256)
        if n < 0:
            _lib.check(n)
        return out[:n]

    def debug_fused(self, flags: int) -> None:
        _lib.check(self._L.rsr_debug_fused(self._h, flags))

    def debug_read_fused_profile(self) -> np.ndarray:
        out = np.zeros((256, 32), np.int64)
        n = self._L.rsr_debug_read_fused_profile(self._h, _as_ptr(out), 256)
        if n < 0:
            _lib.check(n)
        return out[:n]

    def plan(self) -> str:
        """Launch plan of this context: split / fused middle levels, which launches run on CTA pairs."""
        return self._L.rsr_plan(self._h).decode()

    def last_launch_count(self) -> int:
        return int(self._L.rsr_last_launch_count(self._h))


def bilinear(image: np.ndarray, factor: int = FACTOR, device: int = 0) -> np.ndarray:
    """`-p bilinear` = bilinear_net(factor) (network.rs:111-123, main.rs:153-155): interpolation in linear light.
    u8 [H][W][3|4] -> u8 RGBA, or f32 [H][W][3] -> f32 [fH][fW][3]."""
    L = _lib.load()
    a = np.ascontiguousarray(image)
    H, W = a.shape[:2]
    if a.dtype == np.uint8:
        out = np.empty((H * factor, W * factor, 4), np.uint8)
        _lib.check(L.rsr_bilinear_u8(device, _as_ptr(a), W, H, a.shape[2], factor, _as_ptr(out)))
    else:
        a = np.ascontiguousarray(a, dtype=np.float32)
        out = np.empty((H * factor, W * factor, 3), np.float32)
        _lib.check(L.rsr_bilinear_f32(device, _as_ptr(a), W, H, factor, _as_ptr(out)))
    return out


def downsample(image: np.ndarray, factor: int = FACTOR, device: int = 0) -> np.ndarray:
    """`-d` = downsample_net(factor) (network.rs:125-138, main.rs:139-141): average pooling in linear light;
    this is also how the training branch synthesises its low-resolution inputs (network.rs:90-92)."""
    L = _lib.load()
    a = np.ascontiguousarray(image)
    H, W = a.shape[:2]
    oh, ow = -(-H // factor), -(-W // factor)
    if a.dtype == np.uint8:
        out = np.empty((oh, ow, 4), np.uint8)
        _lib.check(L.rsr_downsample_u8(device, _as_ptr(a), W, H, a.shape[2], factor, _as_ptr(out)))
    else:
        a = np.ascontiguousarray(a, dtype=np.float32)
        out = np.empty((oh, ow, 3), np.float32)
        _lib.check(L.rsr_downsample_f32(device, _as_ptr(a), W, H, factor, _as_ptr(out)))
    return out


def conv2d_3x3(x: np.ndarray, weights: np.ndarray, device: int = 0, iters: int = 1):
    """Generic zero-padded 3x3 convolution on the sr_net kernel template (BASELINE config 5 sweep).
    x: f32 [H][W][Cin], weights: f32 [Cout][3][3][Cin] (alumina layout); Cin, Cout multiples of 32.
    Returns (out f32 [H][W][Cout], mean kernel milliseconds over `iters` runs)."""
    L = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    wt = np.ascontiguousarray(weights, dtype=np.float32)
    H, W, ci = x.shape
    co = wt.shape[0]
    assert wt.shape == (co, 3, 3, ci)
    out = np.empty((H, W, co), np.float32)
    ms = ctypes.c_float(0.0)
    _lib.check(L.rsr_conv2d_3x3_f32(device, _as_ptr(x), _as_ptr(wt), W, H, ci, co, _as_ptr(out), iters, ctypes.byref(ms)))
    return out, float(ms.value)



class MultiUpscaler:
    """Several GPUs from one host process (rsr_create_multi, include/rsr_b200.h): one context per listed device, one host
    thread per context inside each call, no collective.  `devices=None`: every visible device."""

    def __init__(self, params: Optional[np.ndarray] = None, parameters: str = "imagenet", factor: int = FACTOR,
                 devices: Optional[Sequence[int]] = None, flags: int = 0):
        if params is None:
            params = load_builtin_params(parameters)
        params = np.ascontiguousarray(params, dtype=np.float32).ravel()
        self._L = _lib.load()
        self.factor = factor
        dev = None if devices is None else (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        h = self._L.rsr_create_multi(params.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), params.size, factor,
                                     dev, 0 if devices is None else len(devices), int(flags))
        if not h:
            code = self._L.rsr_last_error_code()
            msg = self._L.rsr_last_error().decode("utf-8", "replace")
            if code == _lib.RSR_E_PARAM_COUNT:
                raise ValueError(msg)
            raise RsrError(code, msg)
        self._h = ctypes.c_void_p(h)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.rsr_destroy_multi(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def size(self) -> int:
        return int(self._L.rsr_multi_size(self._h))

    def forward_batch(self, x_nhwc3: np.ndarray) -> np.ndarray:
        """Image batch, image i on context i mod G.  f32 [N][H][W][3] -> f32 [N][fH][fW][3]."""
        x = np.ascontiguousarray(x_nhwc3, dtype=np.float32)
        if x.ndim != 4 or x.shape[3] != 3:
            raise ValueError("expected [N][H][W][3]")
        N, H, W = x.shape[:3]
        out = np.empty((N, H * self.factor, W * self.factor, 3), np.float32)
        _lib.check(self._L.rsr_multi_upscale_batch_f32(self._h, _as_ptr(x), N, W, H, _as_ptr(out)))
        return out

    def forward_batch_u8(self, imgs: np.ndarray) -> np.ndarray:
        """u8 [N][H][W][3|4] -> RGBA u8 [N][fH][fW][4]."""
        x = np.ascontiguousarray(imgs, dtype=np.uint8)
        if x.ndim != 4 or x.shape[3] not in (3, 4):
            raise ValueError("expected [N][H][W][3|4]")
        N, H, W, C = x.shape
        out = np.empty((N, H * self.factor, W * self.factor, 4), np.uint8)
        _lib.check(self._L.rsr_multi_upscale_batch_u8(self._h, _as_ptr(x), N, W, H, C, _as_ptr(out)))
        return out

    def forward_frame(self, x_hwc3: np.ndarray) -> np.ndarray:
        """ONE frame as G horizontal bands with 7-row input halos; bit-identical to Upscaler.forward of the frame."""
        x = np.ascontiguousarray(x_hwc3, dtype=np.float32)
        if x.ndim != 3 or x.shape[2] != 3:
            raise ValueError("expected [H][W][3]")
        H, W = x.shape[:2]
        out = np.empty((H * self.factor, W * self.factor, 3), np.float32)
        _lib.check(self._L.rsr_multi_upscale_frame_f32(self._h, _as_ptr(x), W, H, _as_ptr(out)))
        return out

    def forward_frame_u8(self, img: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(img, dtype=np.uint8)
        if x.ndim != 3 or x.shape[2] not in (3, 4):
            raise ValueError("expected [H][W][3|4]")
        H, W, C = x.shape
        out = np.empty((H * self.factor, W * self.factor, 4), np.uint8)
        _lib.check(self._L.rsr_multi_upscale_frame_u8(self._h, _as_ptr(x), W, H, C, _as_ptr(out)))
        return out


def upscale(input_image: np.ndarray, parameters: str = "imagenet", custom: Optional[str] = None,
            device: int = 0) -> np.ndarray:
    """The reference's upscale(input_image) -> output_image (main.rs:130-178) for the neural
    modes: u8 RGB/RGBA image in, u8 RGBA image out (3x).  `custom` is a .rsr path (`-c`)."""
    if custom is not None:
        with open(custom, "rb") as fh:  # "Error opening parameter file" main.rs:134
            params = decode_rsr(fh.read())
    else:
        params = load_builtin_params(parameters)
    with Upscaler(params, device=device) as up:
        return up.forward_u8(input_image)
