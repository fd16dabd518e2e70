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


PREC_FLAGS = R.RSR_FLAG_BF16X3 if "bf16x3" in sys.argv[1:] else 0   # operand format under test (default: fp16 + mxfp8)


def stage_check(params, x, label):
    H, W = x.shape[:2]
    npx = H * W
    ref, taps = O.forward(params, x, 3, np.float64, taps=True)
    up = R.Upscaler(params, flags=R.RSR_FLAG_SPLIT_MID | PREC_FLAGS)   # the split plan leaves every node in HBM
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
    up.set_kernel_timing(True)
    t = time.time()
    out = up.forward(x)
    dt = time.time() - t
    err = np.abs(out - ref)
    lin = np.abs(O.srgb_to_linear(out) - O.srgb_to_linear(ref))
    print(f"[{label}] OUTPUT max|err| sRGB = {err.max():.3e}  linear-light = {lin.max():.3e}  ({dt*1e3:.1f} ms host call)")
    print(f"[{label}] kernel ms: {up.last_kernel_times_ms()}")
    up.close()
    return ok and err.max() < 1e-3


def fused_check(params, x, label):
    """The fused middle launch (conv_fused.cuh, RSR_FLAG_FUSED_MID): l3 and the output against the oracle,
    and the output against the split launch plan."""
    H, W = x.shape[:2]
    ref, taps = O.forward(params, x, 3, np.float64, taps=True)
    try:
        up = R.Upscaler(params, flags=R.RSR_FLAG_FUSED_MID | PREC_FLAGS)
        up.debug_stop_after(3)
        up.forward(x)
        l3 = up.debug_read_node(0, H * W).reshape(H, W, 32)
        err = np.abs(l3 - taps["l3_activ"])
        print(f"[fused {label}] l3_activ  max|err| = {err.max():.3e} worst at {np.unravel_index(err.argmax(), err.shape)}")
        if not err.max() < 1e-3:
            bad = err.max(axis=2) > 1e-3
            ys, xs = np.nonzero(bad)
            print(f"    bad pixels: {bad.sum()} of {bad.size}; y {ys.min()}..{ys.max()} x {xs.min()}..{xs.max()};"
                  f" bad columns mod 124: {sorted(set((xs % 124).tolist()))[:20]}")
        up.debug_stop_after(-1)
        up.set_kernel_timing(True)
        out = up.forward(x)
        launches = up.last_launch_count()
        times = up.last_kernel_times_ms()
        up.close()
        split = R.Upscaler(params, flags=R.RSR_FLAG_SPLIT_MID | PREC_FLAGS)
        out_s = split.forward(x)
        split.close()
    except Exception as e:  # noqa
        print(f"[fused {label}] EXCEPTION {e}")
        return False
    e_out = float(np.abs(out - ref).max())
    e_split = float(np.abs(out - out_s).max())
    print(f"[fused {label}] OUTPUT max|err| vs oracle = {e_out:.3e}, vs split plan = {e_split:.3e}; launches {launches}")
    print(f"[fused {label}] kernel ms: {times}")
    return bool(err.max() < 1e-3 and e_out < 1e-3 and e_split < 1e-4 and launches == 4)


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
    fok = True
    for prm, x, label in ((P["anime"], O.img_to_data(load_png("cartoon_lr")), "cartoon 84x120"),
                          (P["imagenet"], noise_image(40, 300, 5), "noise 300x40"),
                          (P["imagenetlinear"], noise_image(7, 5, 6), "noise 5x7"),
                          (P["imagenet"], noise_image(1, 1, 7), "noise 1x1"),
                          (P["imagenet"], noise_image(300, 125, 8), "noise 125x300"),
                          (P["imagenet"], O.img_to_data(load_png("butterfly_lr")), "butterfly 256x171")):
        fok &= fused_check(prm, x, label)
        if not fok:
            break
    print("FUSED_BRINGUP", "PASS" if fok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
