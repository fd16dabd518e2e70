"""GPU diagnostic for the fused middle launch: compares its l2 ring rows and raw accumulators with the oracle
(test hooks rsr_debug_fused / rsr_debug_read_node 5, 6) and prints the MMA warp's wait-cycle profile."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rusty_sr_b200 as R  # noqa: E402
from conftest import load_png, noise_image  # noqa: E402
from oracle import oracle as O  # noqa: E402


def wconv(params, name):
    sp = {s.name: s for s in R.param_layout()}[name]
    return params[sp.offset:sp.offset + sp.count].reshape(sp.shape)


def report(name, got, want):
    err = np.abs(got - want)
    print(f"  {name:28s} max|err| {err.max():.3e}  (ref max {np.abs(want).max():.3f})", end="")
    if err.max() > 1e-3:
        bad = err.max(axis=2) > 1e-3
        ys, xs = np.nonzero(bad)
        chans = np.nonzero((err > 1e-3).any(axis=(0, 1)))[0]
        print(f"\n      bad px {bad.sum()}/{bad.size}; y {ys.min()}..{ys.max()} x {xs.min()}..{xs.max()};"
              f" bad channels {chans.tolist()[:40]}")
        # is the data a permutation / shift of the expected tensor?
        for dy, dx in ((0, 0), (0, 1), (0, -1), (1, 0), (-1, 0)):
            sh = np.roll(want, (dy, dx), axis=(0, 1))
            e = np.abs(got - sh)[2:-2, 2:-2]
            print(f"      vs expected shifted by (dy={dy}, dx={dx}): interior max|err| {e.max():.3e}")
        print(f"      got[2,2,:8]  = {np.round(got[2, 2, :8], 4)}\n      want[2,2,:8] = {np.round(want[2, 2, :8], 4)}")
    else:
        print()


def check(params, x, label):
    H, W = x.shape[:2]
    ref, taps = O.forward(params, x, 3, np.float64, taps=True)
    l1, l2 = taps["l1_activ"].astype(np.float32), taps["l2_activ"].astype(np.float32)
    c6, c8 = O.conv2d(l1, wconv(params, "conv6")), O.conv2d(l2, wconv(params, "conv8"))
    print(f"[{label}]")
    up = R.Upscaler(params, flags=R.RSR_FLAG_FUSED_MID)
    for mode, want_t3, tag in ((3, c6, "S2 skipped: conv6(l1)"), (2, c6 + c8, "conv6(l1)+conv8(l2)")):
        up.debug_fused(mode)
        up.debug_stop_after(3)
        try:
            up.forward(x)
            got_l2 = up.debug_read_node(5, H * W).reshape(H, W, 32)
            got_t3 = up.debug_read_node(6, H * W).reshape(H, W, 32)
        except Exception as e:  # noqa
            print("  EXCEPTION", e)
            return
        if mode == 3:
            report("l2 (ring rows)", got_l2, taps["l2_activ"])
        report("T[32:64] " + tag, got_t3, want_t3)
    up.debug_fused(0)
    up.debug_stop_after(-1)
    out = up.forward(x)
    print(f"  OUTPUT max|err| {np.abs(out - ref).max():.3e}")
    up.close()


def profile():
    up = R.Upscaler(parameters="imagenet", flags=R.RSR_FLAG_FUSED_MID)
    x = noise_image(1080, 1920, 1234)
    up.forward(x)
    up.debug_fused(4)
    up.forward(x)
    pr = up.debug_read_fused_profile()[:148].astype(np.float64)
    names = ["t_empty", "a_full", "b_full(S1)", "l2_full", "b_full(S2)"]
    tot = pr[:, 7].mean()
    print("fused MMA-warp wait profile (mean cycles per CTA, % of total):", f"total {tot:.0f}")
    for i, n in enumerate(names):
        print(f"   {n:12s} {pr[:, i].mean():10.0f}  {100 * pr[:, i].mean() / tot:5.1f} %")
    print("   kernel ms:", up.last_kernel_times_ms())
    up.close()


if __name__ == "__main__":
    P = {k: R.load_builtin_params(k) for k in R.BUILTIN_PARAMETERS}
    check(P["anime"], O.img_to_data(load_png("cartoon_lr")), "cartoon 84x120")
    check(P["imagenet"], noise_image(40, 300, 5), "noise 300x40")
    if "--profile" in sys.argv:
        profile()
