"""ctypes binding of librsr_b200.so (include/rsr_b200.h).  Fails loudly if the CUDA library is
missing: there is no CPU or PyTorch fallback anywhere in this package."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RSR_B200_LIB") or os.path.join(_HERE, "lib", "librsr_b200.so")  # env override: tools/ablate.sh

RSR_OK = 0
RSR_E_PARAM_COUNT = -1
RSR_E_BAD_ARG = -2
RSR_E_CUDA = -3
RSR_E_OOM = -4
RSR_E_UNSUPPORTED = -5
RSR_FLAG_FUSED_MID = 1   # rsr_create flags (include/rsr_b200.h)
RSR_FLAG_SPLIT_MID = 2
RSR_FLAG_NO_PAIRS = 4    # every launch on single CTAs (no cta_group::2 kernels)
RSR_FLAG_BF16X3 = 8      # three bf16 products per term instead of fp16 + two fp8 corrections (more accurate, slower)

# every symbol include/rsr_b200.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "rsr_num_params", "rsr_create", "rsr_destroy", "rsr_upscale_f32", "rsr_upscale_u8",
    "rsr_upscale_batch_f32", "rsr_upscale_batch_u8", "rsr_upscale_device_f32",
    "rsr_upscale_device_u8", "rsr_synchronize", "rsr_submit_u8", "rsr_submit_f32", "rsr_wait", "rsr_wait_all", "rsr_upscale_band_f32", "rsr_upscale_band_u8", "rsr_set_stream",
    "rsr_create_multi", "rsr_destroy_multi", "rsr_multi_size", "rsr_multi_ctx", "rsr_multi_upscale_batch_f32",
    "rsr_multi_upscale_batch_u8", "rsr_multi_upscale_frame_f32", "rsr_multi_upscale_frame_u8",
    "rsr_last_kernel_times", "rsr_kernel_name", "rsr_last_launch_count", "rsr_plan", "rsr_last_error",
    "rsr_last_error_code", "rsr_version", "rsr_debug_stop_after", "rsr_debug_read_node",
    "rsr_debug_profile", "rsr_debug_read_profile", "rsr_debug_fused", "rsr_debug_read_fused_profile",
    "rsr_bilinear_f32", "rsr_bilinear_u8", "rsr_downsample_f32", "rsr_downsample_u8", "rsr_conv2d_3x3_f32",
    "rsr_upscale_u8_rgb", "rsr_upscale_device_u8_rgb", "rsr_upscale_band_device_f32", "rsr_upscale_band_device_u8",
    "rsr_submit_u8_rgb", "rsr_host_alloc", "rsr_host_free", "rsr_rsr_encoded_size", "rsr_rsr_encode", "rsr_rsr_decode", "rsr_debug_create", "rsr_set_kernel_timing",
)

_lib = None


class RsrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rsr_b200 error {code}: {message}")
        self.code = code


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C rusty_sr_b200/csrc`). rusty_sr_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, fp, u8p, ci = ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p, ctypes.c_int
    L.rsr_num_params.restype = ctypes.c_size_t
    L.rsr_num_params.argtypes = [ci]
    L.rsr_create.restype = vp
    L.rsr_create.argtypes = [fp, ctypes.c_size_t, ci, ci, ci]
    L.rsr_debug_create.restype = vp
    L.rsr_debug_create.argtypes = [fp, ctypes.c_size_t, ci, ci, ci, ci]
    L.rsr_destroy.restype = None
    L.rsr_destroy.argtypes = [vp]
    L.rsr_upscale_f32.argtypes = [vp, vp, ci, ci, vp]
    L.rsr_upscale_u8.argtypes = [vp, u8p, ci, ci, ci, u8p]
    L.rsr_upscale_batch_f32.argtypes = [vp, vp, ci, ci, ci, vp]
    L.rsr_upscale_batch_u8.argtypes = [vp, u8p, ci, ci, ci, ci, u8p]
    L.rsr_upscale_device_f32.argtypes = [vp, vp, ci, ci, ci, vp, vp]
    L.rsr_upscale_device_u8.argtypes = [vp, vp, ci, ci, ci, ci, vp, vp]
    L.rsr_synchronize.argtypes = [vp]
    L.rsr_upscale_u8_rgb.argtypes = [vp, u8p, ci, ci, ci, u8p]
    L.rsr_upscale_device_u8_rgb.argtypes = [vp, vp, ci, ci, ci, ci, vp]
    L.rsr_upscale_band_device_f32.argtypes = [vp, vp, ci, ci, ci, ci, vp]
    L.rsr_upscale_band_device_u8.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, ci]
    L.rsr_submit_u8_rgb.argtypes = [vp, u8p, ci, ci, ci, u8p]
    L.rsr_rsr_encoded_size.restype = ctypes.c_size_t
    L.rsr_rsr_encoded_size.argtypes = [ctypes.c_size_t]
    L.rsr_rsr_encode.restype = ctypes.c_longlong
    L.rsr_rsr_encode.argtypes = [vp, ctypes.c_size_t, vp, ctypes.c_size_t]
    L.rsr_rsr_decode.restype = ctypes.c_longlong
    L.rsr_rsr_decode.argtypes = [vp, ctypes.c_size_t, vp, ctypes.c_size_t]
    L.rsr_submit_u8.argtypes = [vp, u8p, ci, ci, ci, u8p]
    L.rsr_submit_f32.argtypes = [vp, vp, ci, ci, vp]
    L.rsr_wait.argtypes = [vp, ci]
    L.rsr_wait_all.argtypes = [vp]
    L.rsr_upscale_band_f32.argtypes = [vp, vp, ci, ci, ci, ci, vp]
    L.rsr_upscale_band_u8.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp]
    L.rsr_create_multi.restype = vp
    L.rsr_create_multi.argtypes = [vp, ctypes.c_size_t, ci, vp, ci, ci]
    L.rsr_destroy_multi.argtypes = [vp]
    L.rsr_destroy_multi.restype = None
    L.rsr_multi_size.argtypes = [vp]
    L.rsr_multi_ctx.argtypes = [vp, ci]
    L.rsr_multi_ctx.restype = vp
    L.rsr_multi_upscale_batch_f32.argtypes = [vp, vp, ci, ci, ci, vp]
    L.rsr_multi_upscale_batch_u8.argtypes = [vp, vp, ci, ci, ci, ci, vp]
    L.rsr_multi_upscale_frame_f32.argtypes = [vp, vp, ci, ci, vp]
    L.rsr_multi_upscale_frame_u8.argtypes = [vp, vp, ci, ci, ci, vp]
    L.rsr_set_stream.argtypes = [vp, vp]
    L.rsr_last_kernel_times.argtypes = [vp, fp, ci]
    L.rsr_set_kernel_timing.argtypes = [vp, ci]
    L.rsr_set_kernel_timing.restype = ci
    L.rsr_kernel_name.restype = ctypes.c_char_p
    L.rsr_kernel_name.argtypes = [ci]
    L.rsr_last_launch_count.argtypes = [vp]
    L.rsr_plan.argtypes = [vp]
    L.rsr_plan.restype = ctypes.c_char_p
    L.rsr_last_error.restype = ctypes.c_char_p
    L.rsr_host_alloc.restype = ctypes.c_void_p
    L.rsr_host_alloc.argtypes = [ctypes.c_size_t]
    L.rsr_host_free.restype = None
    L.rsr_host_free.argtypes = [ctypes.c_void_p]
    L.rsr_last_error_code.restype = ci
    L.rsr_version.restype = ctypes.c_char_p
    L.rsr_bilinear_f32.argtypes = [ci, vp, ci, ci, ci, vp]
    L.rsr_bilinear_u8.argtypes = [ci, vp, ci, ci, ci, ci, vp]
    L.rsr_downsample_f32.argtypes = [ci, vp, ci, ci, ci, vp]
    L.rsr_downsample_u8.argtypes = [ci, vp, ci, ci, ci, ci, vp]
    for _n in ("rsr_bilinear_f32", "rsr_bilinear_u8", "rsr_downsample_f32", "rsr_downsample_u8"):
        getattr(L, _n).restype = ci
    L.rsr_conv2d_3x3_f32.argtypes = [ci, vp, vp, ci, ci, ci, ci, vp, ci, fp]
    L.rsr_conv2d_3x3_f32.restype = ci
    L.rsr_debug_stop_after.argtypes = [vp, ci]
    L.rsr_debug_stop_after.restype = ci
    L.rsr_debug_read_node.argtypes = [vp, ci, vp, ctypes.c_size_t]
    L.rsr_debug_read_node.restype = ci
    L.rsr_debug_fused.argtypes = [vp, ci]
    L.rsr_debug_fused.restype = ci
    L.rsr_debug_read_fused_profile.argtypes = [vp, vp, ci]
    L.rsr_debug_read_fused_profile.restype = ci
    L.rsr_debug_profile.argtypes = [vp, ci]
    L.rsr_debug_profile.restype = ci
    L.rsr_debug_read_profile.argtypes = [vp, ci, vp, ci]
    L.rsr_debug_read_profile.restype = ci
    for name in ("rsr_upscale_f32", "rsr_upscale_u8", "rsr_upscale_batch_f32", "rsr_upscale_batch_u8",
                 "rsr_upscale_device_f32", "rsr_upscale_device_u8", "rsr_synchronize", "rsr_submit_u8",
                 "rsr_submit_f32", "rsr_wait", "rsr_wait_all", "rsr_submit_u8", "rsr_submit_f32", "rsr_wait", "rsr_wait_all",
                 "rsr_upscale_band_f32", "rsr_upscale_band_u8", "rsr_set_stream", "rsr_last_kernel_times",
                 "rsr_last_launch_count", "rsr_upscale_u8_rgb", "rsr_upscale_device_u8_rgb", "rsr_upscale_band_device_f32",
                 "rsr_upscale_band_device_u8", "rsr_submit_u8_rgb"):
        getattr(L, name).restype = ci
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        L = load()
        raise RsrError(rc, L.rsr_last_error().decode("utf-8", "replace"))
