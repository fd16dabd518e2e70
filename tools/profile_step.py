"""Runs a few 1080p forward passes (device-resident) — the command ncu wraps.
usage: python tools/profile_step.py [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (device buffers only)

import rusty_sr_b200 as R  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
W, H = 1920, 1080
rng = np.random.default_rng(1234)
x = torch.from_numpy(rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8).astype(np.float32) / 255.0).cuda()
out = torch.empty((3 * H, 3 * W, 3), dtype=torch.float32, device="cuda")
up = R.Upscaler(parameters="imagenet")
up.set_kernel_timing(os.environ.get("PROFILE_TIMING", "1") == "1")   # PROFILE_TIMING=0: launches overlap as in production
for _ in range(steps):
    up.forward_device(x.data_ptr(), 1, W, H, d_out_f32=out.data_ptr())
up.synchronize()
print(up.last_kernel_times_ms())
