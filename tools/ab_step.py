"""Per-launch CUDA-event times of a few 1080p forward passes, without importing torch (fast start-up): the unit
of tools/ab_session.py same-box A/B runs.  RSR_B200_LIB selects the library variant, AB_FLAGS the rsr_create flags
(e.g. 8 = RSR_FLAG_BF16X3, 2 = RSR_FLAG_SPLIT_MID), AB_W / AB_H the frame size.
usage: python tools/ab_step.py [steps]   -> one JSON line {"lib": ..., "plan": ..., "us": {launch: mean microseconds}}"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rusty_sr_b200 as R  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
W, H = int(os.environ.get("AB_W", 1920)), int(os.environ.get("AB_H", 1080))
rng = np.random.default_rng(1234)
x = (rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8).astype(np.float32) / 255.0)
up = R.Upscaler(parameters="imagenet", flags=int(os.environ.get("AB_FLAGS", "0")))
up.set_kernel_timing(True)
acc = {}
for i in range(steps + 1):
    up.forward(x)
    if i == 0:
        continue   # warm-up (allocation, first-touch)
    for k, v in up.last_kernel_times_ms().items():
        acc.setdefault(k, []).append(v * 1e3)
print(json.dumps({"lib": os.path.basename(os.environ.get("RSR_B200_LIB", "shipped")), "plan": up.plan(),
                  "us": {k: float(np.mean(v)) for k, v in acc.items()}}))
up.close()
