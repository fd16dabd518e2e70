"""Per-CTA duration spread of the launches (needs a library built with -DRSR_PROFILE=1 -DRSR_FUSED_PROF=1, RSR_B200_LIB=...):
the kernels partition their row tiles statically, so a launch lasts as long as its slowest CTA.  Prints, per profiled launch,
mean / max cycles per CTA and how the duration correlates with the number of work items (strip crossings) of the CTA's range.
usage (under gpurun): RSR_B200_LIB=$PWD/build/abl/new_prof.so python tools/cta_spread.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rusty_sr_b200 as R  # noqa: E402

W, H = 1920, 1080
rng = np.random.default_rng(1234)
x = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8).astype(np.float32) / 255.0


def items(total_rows_per_strip, n_strips, n_ranges):
    """number of work items (strip pieces) of each of the n_ranges equal ranges of the strip-major row-tile list"""
    total = total_rows_per_strip * n_strips
    out = []
    for i in range(n_ranges):
        a, b = total * i // n_ranges, total * (i + 1) // n_ranges
        out.append((b - 1) // total_rows_per_strip - a // total_rows_per_strip + 1)
    return np.array(out)


up = R.Upscaler(parameters="imagenet")
up.set_kernel_timing(True)
up.forward(x)
up.debug_profile(True)
up.forward(x)
kt = up.last_kernel_times_ms()
for k, name in enumerate(list(kt), start=1):
    prof = up.debug_read_profile(k).astype(np.float64)
    if len(prof) == 0:
        continue
    tot = prof[:, 15]
    it = items(H, (W + 127) // 128, len(tot))
    print(f"{name:20s} {kt[name]*1e3:7.1f} us  CTAs {len(tot)}  cycles mean {tot.mean():9.0f} max {tot.max():9.0f} (+{100*(tot.max()/tot.mean()-1):.1f} %) "
          f"std {100*tot.std()/tot.mean():.1f} %;  one-item CTAs {tot[it == 1].mean():9.0f}  two-item CTAs {tot[it == 2].mean() if (it == 2).any() else 0:9.0f}")
    order = np.argsort(tot)
    print("   slowest CTAs:", [(int(i), int(it[i]), round(tot[i] / tot.mean(), 3)) for i in order[-6:]], " fastest:", [(int(i), int(it[i]), round(tot[i] / tot.mean(), 3)) for i in order[:4]])
up.debug_profile(False)
for rep in range(2):
  up.debug_fused(8)
  up.forward(x)
  pr = up.debug_read_fused_profile().astype(np.float64)
  up.debug_fused(0)
  if len(pr):
      tot = pr[0::2, 11] if pr[1::2, 11].sum() == 0 else pr[:, 11]      # FP_MMA_TOTAL: the leader CTA of every pair
      tot = tot[tot > 0]
      it = items(H, (W + 123) // 124, len(tot))
      print(f"fused launch          CTAs(pairs) {len(tot)}  cycles mean {tot.mean():9.0f} max {tot.max():9.0f} (+{100*(tot.max()/tot.mean()-1):.1f} %) std {100*tot.std()/tot.mean():.1f} %;"
            f"  one-item {tot[it == 1].mean():9.0f}  two-item {tot[it == 2].mean() if (it == 2).any() else 0:9.0f}")
      order = np.argsort(tot)
      print("   slowest:", [(int(i), int(it[i]), round(tot[i] / tot.mean(), 3)) for i in order[-6:]], " fastest:", [(int(i), int(it[i]), round(tot[i] / tot.mean(), 3)) for i in order[:4]])
      lead = pr[0::2] if pr[1::2, 11].sum() == 0 else pr
      lead = lead[lead[:, 11] > 0]
      smid, t0, t1 = lead[:, 21], lead[:, 22], lead[:, 23]
      print("   start spread (us): %.1f   end spread (us): %.1f   duration by globaltimer mean %.1f max %.1f min %.1f us" %
            ((t0.max() - t0.min()) / 1e3, (t1.max() - t1.min()) / 1e3, (t1 - t0).mean() / 1e3, (t1 - t0).max() / 1e3, (t1 - t0).min() / 1e3))
      print("   (pair, smid, duration/mean):", [(int(i), int(smid[i]), round((t1[i] - t0[i]) / (t1 - t0).mean(), 3)) for i in range(len(lead))])
