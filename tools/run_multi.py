"""BASELINE configs 3 and 4 through the single-process multi-GPU C ABI (rsr_create_multi; Python MultiUpscaler):
  config 3: batch of 32 synthetic 256x256 u8 patches, image i -> GPU i mod G (rng seed 1234 + i)
  config 4: one 7680x4320 u8 frame (seed 4321) in G horizontal bands with 7-row halos
Host buffers in, host buffers out (u8 RGB -> RGBA u8), so the numbers are END-TO-END (H2D, forward, D2H inside every call)
and timed with the host clock around the synchronous call; kernel-only numbers come from bench.py / tools/run_banded.py.
usage: python tools/run_multi.py [n_gpus | 0 = all] [--check] [--repeat-device]   -> one JSON line per config
  --repeat-device: list device 0 n_gpus times (exercises the sharding logic on a one-GPU box; no speed-up expected)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rusty_sr_b200 as R  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
G = int(args[0]) if args else 0
devices = None if G == 0 else ([0] * G if "--repeat-device" in sys.argv else list(range(G)))
check = "--check" in sys.argv
F = 3


def best_of(fn, n=3):
    fn()   # warm-up: allocates every context's workspace
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), sum(ts) / len(ts)


with R.MultiUpscaler(parameters="imagenet", devices=devices) as m:
    g = m.size()
    # config 3
    patches = np.stack([np.random.default_rng(1234 + i).integers(0, 256, size=(256, 256, 3), dtype=np.uint8) for i in range(32)])
    best, mean = best_of(lambda: m.forward_batch_u8(patches))
    line = {"config": "batch-32 256x256 u8 patches -> 768x768 RGBA, image i -> GPU i mod G", "n_contexts": g, "devices": devices or "all",
            "ms_per_batch_best": best * 1e3, "ms_per_batch_mean": mean * 1e3, "output_mpix_per_s_e2e": 32 * 768 * 768 / 1e6 / best}
    if check:
        with R.Upscaler(parameters="imagenet") as u:
            line["bit_identical_to_one_context"] = bool(np.array_equal(m.forward_batch_u8(patches[:4]), np.stack([u.forward_u8(p) for p in patches[:4]])))
    print(json.dumps(line), flush=True)
    # config 4
    frame = np.random.default_rng(4321).integers(0, 256, size=(4320, 7680, 3), dtype=np.uint8)
    best, mean = best_of(lambda: m.forward_frame_u8(frame), 2)
    line = {"config": f"7680x4320 u8 frame -> 23040x12960 RGBA, {g} band(s) with 7-row halos", "n_contexts": g, "devices": devices or "all",
            "ms_per_frame_best": best * 1e3, "ms_per_frame_mean": mean * 1e3, "output_mpix_per_s_e2e": 7680 * 4320 * F * F / 1e6 / best}
    if check:
        with R.Upscaler(parameters="imagenet") as u:
            line["bit_identical_to_one_context"] = bool(np.array_equal(m.forward_frame_u8(frame), u.forward_u8(frame)))
    print(json.dumps(line), flush=True)
