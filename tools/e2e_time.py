"""Milliseconds per frame of the pipelined host API (rsr_submit_u8_rgb / rsr_wait, 3 frames in flight, pinned buffers, 1080p) over
60-frame runs; RSR_B200_LIB selects a library build.  (A lab build that splits the result copy over two streams / copy engines
measured the same 1.22 ms as the single copy: the copy is bound by PCIe, not by the copy engine.)
usage (under gpurun): python tools/e2e_time.py"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import rusty_sr_b200 as R
up = R.Upscaler(parameters="imagenet")
H, W = 1080, 1920
imgs = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(6, H, W, 3), dtype=np.uint8)).pin_memory().numpy()
out = torch.empty((3, 3 * H, 3 * W, 3), dtype=torch.uint8).pin_memory().numpy()
def run(n):
    t = []
    for i in range(n):
        if i >= 3: up.wait(t[i - 3])
        t.append(up.submit_u8(imgs[i % 6], out[i % 3]))
    up.wait_all()
run(6)
for rep in range(3):
    t0 = time.perf_counter(); run(60); dt = (time.perf_counter() - t0) / 60 * 1e3
    print(os.path.basename(os.environ.get("RSR_B200_LIB", "shipped")), "e2e ms/frame %.4f" % dt)
