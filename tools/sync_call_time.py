"""Milliseconds per SYNCHRONOUS host call (rsr_upscale_u8 / _u8_rgb / _f32, pinned buffers, 1080p): the library runs a large frame
as up to four bands whose result rows travel to the host while the next band computes; with per-launch timing on it runs the
frame in one pass and copies afterwards (the behaviour before).  usage (under gpurun): python tools/sync_call_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, ctypes, numpy as np, torch, rusty_sr_b200 as R
L = R._lib.load()
up = R.Upscaler(parameters="imagenet")
img = np.random.default_rng(0).integers(0,256,size=(1080,1920,3),dtype=np.uint8)
pin_in = torch.from_numpy(img).pin_memory().numpy()
def run(fn, *a):
    R._lib.check(fn(*a)); R._lib.check(fn(*a))
    t=time.perf_counter()
    for i in range(10): R._lib.check(fn(*a))
    return (time.perf_counter()-t)/10*1e3
o4 = torch.empty((3240,5760,4), dtype=torch.uint8).pin_memory().numpy()
o3 = torch.empty((3240,5760,3), dtype=torch.uint8).pin_memory().numpy()
of = torch.empty((3240,5760,3), dtype=torch.float32).pin_memory().numpy()
xf = torch.from_numpy(img.astype(np.float32)/255).pin_memory().numpy()
vp = ctypes.c_void_p
for timing in (False, True):
    up.set_kernel_timing(timing)
    print("whole frame (timing on)" if timing else "banded", 
      "rgba %.3f" % run(L.rsr_upscale_u8, up._h, vp(pin_in.ctypes.data), 1920, 1080, 3, vp(o4.ctypes.data)),
      "rgb %.3f" % run(L.rsr_upscale_u8_rgb, up._h, vp(pin_in.ctypes.data), 1920, 1080, 3, vp(o3.ctypes.data)),
      "f32 %.3f ms" % run(L.rsr_upscale_f32, up._h, xf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 1920, 1080, of.ctypes.data_as(ctypes.POINTER(ctypes.c_float))))
# pageable host buffers (a plain Vec<u8> / numpy array): the copies block the calling thread, the bands behind them still compute
up.set_kernel_timing(False)
pg_in, pg4 = np.array(img), np.empty((3240, 5760, 4), np.uint8)
banded = run(L.rsr_upscale_u8, up._h, vp(pg_in.ctypes.data), 1920, 1080, 3, vp(pg4.ctypes.data))
up.set_kernel_timing(True)
whole = run(L.rsr_upscale_u8, up._h, vp(pg_in.ctypes.data), 1920, 1080, 3, vp(pg4.ctypes.data))
print("pageable buffers: rgba banded %.3f  whole frame %.3f ms" % (banded, whole))

