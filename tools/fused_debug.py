"""GPU diagnostic for the fused middle launch: compares its l2 ring rows and raw accumulators with the oracle
(test hooks rsr_debug_fused / rsr_debug_read_node 5, 6) and prints the wait-cycle profile of every warp role (--profile / --profile-plan / --profile-only)."""
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


FP_NAMES = ["a_empty(l1 producer)", "b_empty(weight producer)", "p2_empty(P2 producer)", "p3_empty(P3 producer)",
            "t_empty(mma)", "a_full(mma)", "b_full S1(mma)", "l2_full(mma)", "b_full S2(mma)", "ISSUE S1(mma)",
            "ISSUE S2(mma)", "TOTAL(mma)", "t1_full(E1)", "p2_full(E1)", "l2_empty(E1)", "TOTAL(E1)", "t2_full(E2)",
            "p3_full(E2)", "store_read(E2)", "TOTAL(E2)", "S1 rows"]


def profile(mode=4):
    """mode 4: the single-CTA debug kernel (any build).  mode 8: the plan's own kernel, single CTA or CTA pair
    (RSR_B200_PAIR bit 3), in a library built with -DRSR_FUSED_PROF=1 (tools/ablate.sh; RSR_B200_LIB selects it)."""
    up = R.Upscaler(parameters="imagenet", flags=R.RSR_FLAG_FUSED_MID | int(os.environ.get("AB_FLAGS", "0")))
    x = noise_image(1080, 1920, 1234)
    up.set_kernel_timing(True)
    up.forward(x)
    up.debug_fused(mode)
    up.forward(x)
    pr = up.debug_read_fused_profile().astype(np.float64)
    # pair plan: only the leader CTA of a cluster runs the MMA role; every CTA runs the others
    mma_rows = pr[pr[:, 11] > 0]
    all_rows = pr[pr[:, 15] > 0]
    if not len(mma_rows):
        print("fused profile: no counters (library built without -DRSR_FUSED_PROF=1?)", up.plan())
        up.close()
        return
    tot = mma_rows[:, 11].mean()
    rows = mma_rows[:, 20].mean()
    print(f"fused wait profile, mode {mode}, plan: {up.plan()}")
    print(f"  CTAs {len(all_rows)} (MMA role in {len(mma_rows)}); MMA-role cycles/CTA {tot:.0f}; S1 rows/CTA {rows:.1f};"
          f" cycles per S1 row {tot / rows:.0f}")
    for i, n in enumerate(FP_NAMES[:20]):
        src = mma_rows if 4 <= i <= 11 else all_rows
        print(f"   {n:26s} {src[:, i].mean():10.0f}  {100 * src[:, i].mean() / tot:5.1f} %   (max CTA {src[:, i].max():10.0f})")
    print("   kernel ms:", up.last_kernel_times_ms())
    up.close()


if __name__ == "__main__":
    if "--profile-only" in sys.argv:
        profile(8)
        sys.exit(0)
    P = {k: R.load_builtin_params(k) for k in R.BUILTIN_PARAMETERS}
    check(P["anime"], O.img_to_data(load_png("cartoon_lr")), "cartoon 84x120")
    check(P["imagenet"], noise_image(40, 300, 5), "noise 300x40")
    if "--profile" in sys.argv:
        profile(4)
    if "--profile-plan" in sys.argv:
        profile(8)
