"""Whole-frame device time of a library variant (RSR_B200_LIB), 1080p device-resident f32 -> f32, CUDA events around
`steps` back-to-back forward passes (no per-launch events: the launches overlap as in production).
usage: python tools/ab_frame.py [steps]  -> one JSON line {"lib", "ms_per_frame"}"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (device buffers and events only)

import rusty_sr_b200 as R  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
W, H = int(os.environ.get("AB_W", 1920)), int(os.environ.get("AB_H", 1080))
rng = np.random.default_rng(1234)
x = torch.from_numpy(rng.integers(0, 256, size=(4, H, W, 3), dtype=np.uint8).astype(np.float32) / 255.0).cuda()
out = torch.empty((3 * H, 3 * W, 3), dtype=torch.float32, device="cuda")
up = R.Upscaler(parameters="imagenet", flags=int(os.environ.get("AB_FLAGS", "0")))
stream = torch.cuda.Stream()
up.set_stream(stream.cuda_stream)
fe = H * W * 3 * 4
for i in range(5):
    up.forward_device(x.data_ptr() + (i % 4) * fe, 1, W, H, d_out_f32=out.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for i in range(steps):
    up.forward_device(x.data_ptr() + (i % 4) * fe, 1, W, H, d_out_f32=out.data_ptr())
e1.record(stream)
torch.cuda.synchronize()
print(json.dumps({"lib": os.path.basename(os.environ.get("RSR_B200_LIB", "shipped")), "ms_per_frame": e0.elapsed_time(e1) / steps}))
up.close()
