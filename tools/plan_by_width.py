"""Fused vs split plan of the middle levels on narrow images: a batch of 32 patches S x S (u8 RGB -> u8 RGB, device resident),
milliseconds per batch with RSR_FLAG_FUSED_MID, RSR_FLAG_SPLIT_MID and the library's own choice (rsr_b200.cu use_fused).
usage (under gpurun): python tools/plan_by_width.py [S ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (device buffers and events only)

import rusty_sr_b200 as R  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [248, 256, 372, 384, 500]
N = 32
for S in sizes:
    d_in = torch.from_numpy(np.random.default_rng(1).integers(0, 256, size=(N, S, S, 3), dtype=np.uint8)).cuda()
    d_out = torch.empty((N, 3 * S, 3 * S, 3), dtype=torch.uint8, device="cuda")
    res = {}
    for name, flags in (("fused", R.RSR_FLAG_FUSED_MID), ("split", R.RSR_FLAG_SPLIT_MID), ("default", 0)):
        with R.Upscaler(parameters="imagenet", flags=flags) as up:
            stream = torch.cuda.Stream()
            up.set_stream(stream.cuda_stream)
            run = lambda: up.forward_device_u8_rgb(d_in.data_ptr(), N, S, S, 3, d_out.data_ptr())
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(20):
                run()
            e1.record(stream)
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 20
            plan = up.plan().split(":")[0]
        res[name + "_plan"] = plan
    print(f"S = {S}: fused {res['fused']:.4f} ms  split {res['split']:.4f} ms  default {res['default']:.4f} ms ({res['default_plan']})", flush=True)
