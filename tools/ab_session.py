"""One GPU session under a hard time budget, driven by a JSON spec (tools/specs/*.json): the whole GPU suite on the shipped
library, bench, hardware probes, an interleaved same-box A/B of library variants (build flags) and launch plans
(environment), a combined library of the winners that must pass the GPU suite and is benched against the shipped one,
then the ncu launch list and one full capture.  Everything lands in gpurun_out/; nothing here changes the shipped library.

  python tools/ab_session.py SPEC --build            here (no GPU): compile the spec's variants into build/abl/
  gpurun -- python tools/ab_session.py SPEC BUDGET   on the GPU box     (DRY=1: control flow only, no GPU commands)

Spec: {"entries": {name: {"flags": "<nvcc -D flags>" | "lib": "<file in build/abl>", "env": {...}, "launches": [...],
       "group": "<variants of one group compete; the best one above `min_gain` joins the combined library>"}},
       "probes": [...tests/cuda binaries], "pytest_env": {...}, "ncu_kernel": "regex", "ncu_skip": n, "min_gain": 0.02}
The entry "base" (no flags) is the reference of every comparison."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
ABL = os.path.join(ROOT, "build", "abl")
SPEC = json.load(open(sys.argv[1]))
BUILD_ONLY = "--build" in sys.argv[2:]
BUDGET = float(sys.argv[2]) if len(sys.argv) > 2 and not BUILD_ONLY else 330.0
T0 = time.time()
DRY = os.environ.get("DRY") == "1"
summary = {"stages": {}, "spec": os.path.basename(sys.argv[1])}
NVCC = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
        "-shared"]
# name -> (library file in build/abl, extra environment, launches whose summed time is compared with `base`, group)
ENTRIES = {"base": ("base.so", {}, None, None)}
for _n, _e in SPEC["entries"].items():
    ENTRIES[_n] = (_e.get("lib", _n + ".so"), _e.get("env", {}), tuple(_e["launches"]) if _e.get("launches") else None, _e.get("group"))
MIN_GAIN = float(SPEC.get("min_gain", 0.02))


def build_variants():
    """Compiles base.so and every entry that has build flags (parallel nvcc; no GPU needed)."""
    os.makedirs(ABL, exist_ok=True)
    jobs = {"base": ""}
    jobs.update({n: e["flags"] for n, e in SPEC["entries"].items() if "flags" in e})
    procs = {}
    for n, flags in jobs.items():
        with open(os.path.join(ABL, n + ".flags"), "w") as f:
            f.write(flags)
        procs[n] = subprocess.Popen([*NVCC, *flags.split(), "-o", os.path.join(ABL, n + ".so"), "rusty_sr_b200/csrc/rsr_b200.cu", "-lcudart"],
                                    cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    ok = True
    for n, p in procs.items():
        out, _ = p.communicate()
        print(f"build {n:16s} rc={p.returncode}", flush=True)
        if p.returncode != 0:
            ok = False
            print(out[-2000:])
    return ok


def left():
    return BUDGET - (time.time() - T0)


def save():
    summary["elapsed_s"] = round(time.time() - T0, 1)
    with open(os.path.join(OUT, "ab_session_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


def run(name, cmd, timeout, env=None, out=None, quiet=False):
    t = min(timeout, left() - 5)
    if t < 8:
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
    if not quiet:
        print(f"[{name}] rc={rc} {time.time() - t1:.1f}s (left {left():.0f}s)", flush=True)
    save()
    return rc, so


def pytest_gpu(tag, lib=None, extra_env=None):
    env = dict(SPEC.get("pytest_env", {}))
    env.update(extra_env or {})
    if lib:
        env["RSR_B200_LIB"] = lib
    rc, so = run(f"pytest -m gpu ({tag})", [sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-q"], 150, env, f"pytest_gpu_{tag}.log")
    print("   ", (so.strip().splitlines() or ["-"])[-1], flush=True)
    return rc


def bench(tag, lib=None, extra_env=None):
    env = dict(extra_env or {})
    if lib:
        env["RSR_B200_LIB"] = lib
    rc, so = run(f"bench {tag}", [sys.executable, "bench.py", "--steps", "20", "--warmup", "3"], 90, env, f"bench_{tag}.json")
    if DRY:
        return {"ms_per_step": 1.5 if lib is None else 1.45}
    try:
        d = json.loads(so.strip().splitlines()[-1])
        print(f"    ms/step {d['ms_per_step']:.4f}  value {d['value']:.1f}  e2e {d['e2e']['value']:.1f}  "
              f"{ {k: round(v * 1e3, 1) for k, v in d['kernel_ms'].items()} }  {d['clocks']['reasons']}", flush=True)
        return d
    except Exception as ex:  # noqa
        print("    bench parse failed:", ex, flush=True)
        return None


def ab_step(name):
    lib, env, _, _ = ENTRIES[name]
    path = os.path.join(ABL, lib)
    if not os.path.exists(path) and not DRY:
        return None
    if DRY:
        import random
        us = {"conv0": 88.0, "conv1|2|3": 560.0, "conv5|6|7": 300.0, "conv8|9": 270.0, "conv10+expand+skip": 213.0}
        if name == os.environ.get("DRY_WIN"):   # pretend this entry wins on its launches (exercises the combined-library path)
            for k in ENTRIES[name][2] or ():
                us[k] *= 0.9
        return {k: v * (1 + random.uniform(-0.004, 0.004)) for k, v in us.items()}
    e = {"RSR_B200_LIB": path}
    e.update(env)
    rc, so = run(f"ab {name}", [sys.executable, "tools/ab_step.py", "6"], 40, e, quiet=True)
    try:
        return json.loads(so.strip().splitlines()[-1])["us"]
    except Exception:  # noqa
        print(f"    ab {name}: no result (rc {rc})", flush=True)
        return None


def main():
    os.makedirs(OUT, exist_ok=True)
    py = sys.executable
    # 1. gate: the whole GPU suite on the shipped library (spec "pytest_env": e.g. RSR_TEST_EXPERIMENTAL=1)
    summary["pytest_shipped_rc"] = pytest_gpu("shipped")
    base1 = bench("shipped")
    # 2. hardware probes (tests/cuda)
    for probe in SPEC.get("probes", []):
        exe = os.path.join(ROOT, "tests", "cuda", probe)
        if os.path.exists(exe) or DRY:
            run(probe, [exe], 30, out=probe + ".txt")
    # 3. same-box A/B, interleaved
    times = {n: [] for n in ENTRIES}
    for rnd in range(2):
        if rnd > 0 and left() < 170:
            break
        for n in ENTRIES:
            if left() < 130:
                break
            us = ab_step(n)
            if us:
                times[n].append(us)
    ab, winners = {}, {}
    if times["base"]:
        nb = len(times["base"])
        bmean = {k: sum(t[k] for t in times["base"]) / nb for k in times["base"][0]}
        print("    base us:", {k: round(v, 1) for k, v in bmean.items()}, f"total {sum(bmean.values()):.1f}", flush=True)
        for n, (lib, env, keys, group) in ENTRIES.items():
            if keys is None or not times[n]:
                continue
            v = sum(sum(t[k] for k in keys) for t in times[n]) / len(times[n])
            b = sum(bmean[k] for k in keys)
            gain = 1.0 - v / b
            ab[n] = {"launches": "+".join(keys), "us": round(v, 1), "base_us": round(b, 1), "gain": round(gain, 4), "runs": len(times[n])}
            print(f"    {n:12s} {'+'.join(keys):20s} {v:7.1f} us vs {b:7.1f}  gain {100 * gain:+.1f} %  ({len(times[n])} runs)", flush=True)
            if group and gain > MIN_GAIN and (group not in winners or gain > winners[group][1]):
                winners[group] = (n, gain)
    summary["ab"] = ab
    summary["winners"] = {k: v[0] for k, v in winners.items()}
    save()
    # 4. winners combined into one library: GPU suite + bench against the shipped one
    if winners and left() > 100:
        flags, cenv = [], {}
        for k, (n, _) in winners.items():
            flags += SPEC["entries"][n].get("flags", "").split()
            cenv.update(SPEC["entries"][n].get("env", {}))
        combo = os.path.join(ABL, "combo.so")
        summary["combo_env"] = cenv
        nv = [*NVCC, *flags, "-o", combo, "rusty_sr_b200/csrc/rsr_b200.cu", "-lcudart"]
        summary["combo_flags"] = flags
        rc, _ = run("build combo", nv, 120)
        if rc == 0 or DRY:
            summary["combo_tests_rc"] = pytest_gpu("combo", combo, cenv)
            if summary["combo_tests_rc"] == 0:
                c1 = bench("combo", combo, cenv)
                base2 = bench("shipped2") if left() > 70 else None
                c2 = bench("combo2", combo, cenv) if left() > 70 else None
                bs = [b["ms_per_step"] for b in (base1, base2) if b]
                cs = [c["ms_per_step"] for c in (c1, c2) if c]
                if bs and cs:
                    summary["combo_vs_shipped"] = {"shipped_ms": bs, "combo_ms": cs, "gain": round(1 - (sum(cs) / len(cs)) / (sum(bs) / len(bs)), 4)}
                    print("    combo vs shipped:", summary["combo_vs_shipped"], flush=True)
    save()
    # 5. profiler passes on the shipped library (numbers under a profiler are never bench values)
    ncu = "ncu"
    run("ncu launch list", [ncu, "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-c", "40", "--csv", "--log-file",
                            os.path.join(OUT, "launches.csv"), py, "tools/profile_step.py", "3"], 60, out="ncu_launches.log")
    if SPEC.get("ncu_kernel"):
        run("ncu full", [ncu, "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:" + SPEC["ncu_kernel"], "-s",
                         str(SPEC.get("ncu_skip", 0)), "-c", "1", "-o", os.path.join(OUT, "prof_session"), "-f", py, "tools/profile_step.py", "3"],
            90, SPEC.get("ncu_env"), out="ncu_full.log")
    save()
    print(json.dumps(summary), flush=True)


if __name__ == "__main__":
    if BUILD_ONLY:
        sys.exit(0 if build_variants() else 1)
    main()
