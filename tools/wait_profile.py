"""Per-role wait-cycle breakdown of the conv launches at 1080p (test hook rsr_debug_profile).
Shows, per launch, the share of the kernel's cycles each warp role spent blocked on each mbarrier."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rusty_sr_b200 as R  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
rng = np.random.default_rng(1234)
x = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8).astype(np.float32) / 255.0
up = R.Upscaler(parameters="imagenet")
up.set_kernel_timing(True)
up.forward(x)
up.debug_profile(True)
up.forward(x)
names = ["a_empty(rowprod)", "t_empty(mma)", "a_full(mma)", "b_full(mma)", "p_full(mma)", "p_empty(pinprod)",
         "b_empty(wprod)", "t_full(epi)", "store_read(epi)", "MMA_ISSUE_BLOCK(mma)", "EPI_BUSY(epi)", "p_full(epi, inside EPI_BUSY)", "tmem_ld(epi)", "out_path(epi: math+stage+store issue)"]
kt = up.last_kernel_times_ms()
for k, kname in enumerate(list(kt), start=1):   # launch index = 1 + position in the kernel-time list
    prof = up.debug_read_profile(k).astype(np.float64)
    if len(prof) == 0:          # CTA-pair launches / the fused launch carry no strip-kernel profile
        continue
    tot = prof[:, 15]
    line = f"{kname:20s} {kt[kname]*1e3:7.1f} us  grid {len(prof):3d}  cycles/CTA mean {tot.mean():9.0f} max {tot.max():9.0f} | "
    line += "  ".join(f"{n} {100*prof[:, i].mean()/tot.mean():4.1f}%" for i, n in enumerate(names))
    print(line)
