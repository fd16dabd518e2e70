"""One GPU session, run under gpurun with a hard time budget: parity tests, bench, wait-cycle profiles, then a same-box
A/B of the library variants in build/abl/ (tools/ab_build.sh) — variants that win on their launch are combined into
one library on the box, which must pass the whole GPU test suite and beat the shipped library in bench.py before it
is reported as a candidate default.  Everything lands in gpurun_out/; nothing here changes the shipped library.
usage: python tools/final_round.py [budget_seconds]   (DRY=1: stub the GPU commands, to check the control flow)"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
ABL = os.path.join(ROOT, "build", "abl")
BUDGET = float(sys.argv[1]) if len(sys.argv) > 1 else 420.0
T0 = time.time()
DRY = os.environ.get("DRY") == "1"
LAUNCH_OF = {"c0": "conv0", "c2": "conv5|6|7", "c3": "conv8|9", "c4": "conv10+expand+skip"}   # variant name prefix -> launch it changes
summary = {"stages": {}}


def left():
    return BUDGET - (time.time() - T0)


def save():
    summary["elapsed_s"] = round(time.time() - T0, 1)
    with open(os.path.join(OUT, "final_round_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


def run(name, cmd, timeout, env=None, out=None):
    """Runs cmd (list) with a timeout clipped to the remaining budget; returns (rc, stdout)."""
    t = min(timeout, left() - 5)
    if t < 10:
        print(f"[{name}] skipped: {left():.0f} s left", flush=True)
        summary["stages"][name] = "skipped (budget)"
        return None, ""
    e = dict(os.environ)
    e.update(env or {})
    t1 = time.time()
    if DRY:
        rc, so, se = 0, "", ""
    else:
        try:
            r = subprocess.run(cmd, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=t)
            rc, so, se = r.returncode, r.stdout, r.stderr
        except subprocess.TimeoutExpired as ex:
            rc = 124
            so = ex.stdout.decode() if isinstance(ex.stdout, bytes) else (ex.stdout or "")
            se = ex.stderr.decode() if isinstance(ex.stderr, bytes) else (ex.stderr or "")
    if out:
        with open(os.path.join(OUT, out), "w") as f:
            f.write(so)
        if se.strip():
            with open(os.path.join(OUT, out + ".err"), "w") as f:
                f.write(se)
    summary["stages"][name] = {"rc": rc, "s": round(time.time() - t1, 1)}
    print(f"[{name}] rc={rc} {time.time() - t1:.1f}s (left {left():.0f}s)", flush=True)
    save()
    return rc, so


def bench(tag, lib=None):
    env = {"RSR_B200_LIB": lib} if lib else {}
    rc, so = run(f"bench {tag}", [sys.executable, "bench.py", "--steps", "20", "--warmup", "3"], 120, env, f"bench_{tag}.json")
    if DRY:
        return {"ms_per_step": 1.5 if lib is None else 1.45, "value": 1.0, "e2e": {"value": 1.0}, "kernel_ms": {}}
    try:
        d = json.loads(so.strip().splitlines()[-1])
        print(f"    ms/step {d['ms_per_step']:.4f}  value {d['value']:.1f}  e2e {d['e2e']['value']:.1f}  "
              f"{ {k: round(v * 1e3, 1) for k, v in d['kernel_ms'].items()} }  {d['clocks']['reasons']}", flush=True)
        return d
    except Exception as ex:  # noqa
        print("    bench parse failed:", ex, flush=True)
        return None


def ab_step(lib):
    if DRY:
        import random
        base = {"conv5|6|7": 300.0, "conv8|9": 270.0, "conv10+expand+skip": 213.0, "conv1|2|3": 560.0, "conv0": 100.0, "pack_input": 52.0}
        name = os.path.basename(lib)[:2]
        if name in LAUNCH_OF and "r4" in lib:
            base[LAUNCH_OF[name]] *= 0.95
        return {k: v * (1 + random.uniform(-0.004, 0.004)) for k, v in base.items()}
    rc, so = run(f"ab {os.path.basename(lib)}", [sys.executable, "tools/ab_step.py", "6"], 40, {"RSR_B200_LIB": lib})
    try:
        return json.loads(so.strip().splitlines()[-1])["us"]
    except Exception:  # noqa
        return None


def main():
    os.makedirs(OUT, exist_ok=True)
    py = sys.executable
    # 1. the gate: every GPU test on the shipped library
    rc, so = run("pytest -m gpu (shipped)", [py, "-m", "pytest", "tests", "-m", "gpu", "-x", "-q"], 240, out="pytest_gpu.log")
    print("   ", (so.strip().splitlines() or ["-"])[-1], flush=True)
    # 2. the headline line
    base1 = bench("shipped")
    # 3. wait-cycle profiles (profiling build; numbers are relative shares, not timings)
    prof = os.path.join(ABL, "librsr_b200_prof.so")
    if os.path.exists(prof):
        run("wait profile split plan, no pairs", [py, "tools/wait_profile.py"], 60, {"RSR_B200_LIB": prof, "RSR_B200_PAIR": "0"},
            "wait_profile_split.txt")
        run("fused profile single CTA", [py, "tools/fused_debug.py", "--profile-only"], 60,
            {"RSR_B200_LIB": prof, "RSR_B200_PAIR": "4"}, "fused_profile_single.txt")
        run("fused profile CTA pairs", [py, "tools/fused_debug.py", "--profile-only"], 60,
            {"RSR_B200_LIB": prof, "RSR_B200_PAIR": "12"}, "fused_profile_pair.txt")
    # 4. same-box A/B of the geometry variants, interleaved
    libs = sorted(f for f in os.listdir(ABL) if f.endswith(".so") and f[:2] in LAUNCH_OF) if os.path.isdir(ABL) else []
    base_lib = os.path.join(ABL, "base.so")
    times = {n: [] for n in ["base.so"] + libs}
    for rnd in range(2):
        if left() < 200 and rnd > 0:   # keep time for the combined-library check
            break
        for n in ["base.so"] + libs:
            if left() < 150:
                break
            us = ab_step(os.path.join(ABL, n))
            if us:
                times[n].append(us)
    ab = {}
    winners = {}
    if times["base.so"]:
        bmean = {k: sum(t[k] for t in times["base.so"]) / len(times["base.so"]) for k in times["base.so"][0]}
        print("    base us:", {k: round(v, 1) for k, v in bmean.items()}, flush=True)
        for n in libs:
            if not times[n]:
                continue
            launch = LAUNCH_OF[n[:2]]
            v = sum(t[launch] for t in times[n]) / len(times[n])
            gain = 1.0 - v / bmean[launch]
            ab[n] = {"launch": launch, "us": round(v, 1), "base_us": round(bmean[launch], 1), "gain": round(gain, 4), "runs": len(times[n])}
            print(f"    {n:14s} {launch:20s} {v:7.1f} us vs {bmean[launch]:7.1f}  gain {100 * gain:+.1f} %  ({len(times[n])} runs)", flush=True)
            if gain > 0.02 and (n[:2] not in winners or gain > winners[n[:2]][1]):
                winners[n[:2]] = (n, gain)
    summary["ab"] = ab
    summary["winners"] = {k: v[0] for k, v in winners.items()}
    save()
    # 5. winners combined into one library: must pass the GPU suite and beat the shipped library in bench.py
    if winners and left() > 120:
        flags = []
        for k, (n, _) in winners.items():
            log = os.path.join(ABL, n[:-3] + ".flags")
            flags += open(log).read().split()
        combo = os.path.join(ABL, "combo.so")
        nv = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", *flags, "-o", combo, "rusty_sr_b200/csrc/rsr_b200.cu", "-lcudart"]
        summary["combo_flags"] = flags
        rc, _ = run("build combo", nv, 150)
        if rc == 0 or DRY:
            rc, so = run("pytest -m gpu (combo)", [py, "-m", "pytest", "tests", "-m", "gpu", "-x", "-q"], 200, {"RSR_B200_LIB": combo},
                         "pytest_gpu_combo.log")
            print("   ", (so.strip().splitlines() or ["-"])[-1], flush=True)
            summary["combo_tests_rc"] = rc
            if rc == 0:
                c1 = bench("combo", combo)
                base2 = bench("shipped2") if left() > 60 else None
                c2 = bench("combo2", combo) if left() > 60 else None
                bs = [b["ms_per_step"] for b in (base1, base2) if b]
                cs = [c["ms_per_step"] for c in (c1, c2) if c]
                if bs and cs:
                    summary["combo_vs_shipped"] = {"shipped_ms": bs, "combo_ms": cs, "gain": round(1 - (sum(cs) / len(cs)) / (sum(bs) / len(bs)), 4)}
                    print("    combo vs shipped:", summary["combo_vs_shipped"], flush=True)
    save()
    print(json.dumps(summary), flush=True)


if __name__ == "__main__":
    main()
