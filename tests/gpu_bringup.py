"""Bring-up diagnostic (run under gpurun, not a pytest): runs the forward pass launch by launch
and compares every intermediate node with the CPU oracle's taps.  Prints max-abs errors."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import rusty_sr_b200 as R
from conftest import load_png, noise_image
from oracle import oracle as O


def stage_check(params, x, label):
    H, W = x.shape[:2]
    npx = H * W
    ref, taps = O.forward(params, x, 3, np.float64, taps=True)
    up = R.Upscaler(params)
    # (stop_after, node buffer, oracle tensor)
    l3_part = None
    plan = [
        (1, 0, taps["f_activ"], "f_activ"),
        (2, 1, taps["l1_activ"], "l1_activ"),
        (2, 2, taps["l2_part_f"], "conv2(f)"),
        (2, 3, taps["l3_part_f"], "conv3(f)"),
        (3, 0, taps["l2_activ"], "l2_activ"),
        (4, 1, taps["l3_activ"], "l3_activ"),
    ]
    ok = True
    for stop, which, want, name in plan:
        up.debug_stop_after(stop)
        try:
            up.forward(x)
            got = up.debug_read_node(which, npx).reshape(H, W, 32)
        except Exception as e:  # noqa
            print(f"[{label}] {name}: EXCEPTION {e}")
            return False
        err = np.abs(got - want)
        print(f"[{label}] {name:10s} max|err| = {err.max():.3e}  (ref range {want.min():.3f}..{want.max():.3f})"
              f" worst at {np.unravel_index(err.argmax(), err.shape)}")
        if not (err.max() < 1e-3):
            ok = False
            # locate pattern of error
            bad = err.max(axis=2) > 1e-3
            ys, xs = np.nonzero(bad)
            print(f"    bad pixels: {bad.sum()} of {bad.size}; y range {ys.min()}..{ys.max()} x range {xs.min()}..{xs.max()}")
            break
    up.debug_stop_after(-1)
    t = time.time()
    out = up.forward(x)
    dt = time.time() - t
    err = np.abs(out - ref)
    lin = np.abs(O.srgb_to_linear(out) - O.srgb_to_linear(ref))
    print(f"[{label}] OUTPUT max|err| sRGB = {err.max():.3e}  linear-light = {lin.max():.3e}  ({dt*1e3:.1f} ms host call)")
    print(f"[{label}] kernel ms: {up.last_kernel_times_ms()}")
    up.close()
    return ok and err.max() < 1e-3


def main():
    P = {k: R.load_builtin_params(k) for k in R.BUILTIN_PARAMETERS}
    print(R._lib.load().rsr_version().decode())
    ok = True
    ok &= stage_check(P["anime"], O.img_to_data(load_png("cartoon_lr")), "cartoon 84x120")
    ok &= stage_check(P["imagenet"], noise_image(40, 300, 5), "noise 300x40")
    ok &= stage_check(P["imagenet"], O.img_to_data(load_png("butterfly_lr")), "butterfly 256x171")
    ok &= stage_check(P["imagenetlinear"], noise_image(7, 5, 6), "noise 5x7")
    ok &= stage_check(P["imagenet"], noise_image(1, 1, 7), "noise 1x1")
    print("BRINGUP", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
