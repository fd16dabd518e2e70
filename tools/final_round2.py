"""Second GPU session of the round (run under gpurun, hard time budget): the whole GPU suite on the shipped library and
on the previous default (fallback), bench, the TMEM-operand probe, a same-box A/B of library variants and launch
plans, a combined library of the winners (GPU suite + bench), then the ncu launch list and one full capture.
Everything lands in gpurun_out/.  usage: python tools/final_round2.py [budget_seconds]   (DRY=1: control flow only)"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
ABL = os.path.join(ROOT, "build", "abl")
BUDGET = float(sys.argv[1]) if len(sys.argv) > 1 else 330.0
T0 = time.time()
DRY = os.environ.get("DRY") == "1"
summary = {"stages": {}}
MID = ("conv5|6|7", "conv8|9")
# name -> (library file in build/abl, extra environment, launches whose summed time is compared with `base`)
ENTRIES = {
    "base": ("base.so", {}, None),
    "c0_r4n12": ("c0_r4n12.so", {}, ("conv0",)),
    "c0_r8n12": ("c0_r8n12.so", {}, ("conv0",)),
    "c0_r8n20": ("c0_r8n20.so", {}, ("conv0",)),
    "c0_r6n14": ("c0_r6n14.so", {}, ("conv0",)),
    "c3_r3": ("c3_r3.so", {}, ("conv8|9",)),
    "c3_r3b": ("c3_r3b.so", {}, ("conv8|9",)),
    "pk_v2": ("pk_v2.so", {}, ("pack_input",)),
    "f_single": ("base.so", {"RSR_B200_FUSED": "1", "RSR_B200_PAIR": "4"}, MID),
    "f_pair": ("base.so", {"RSR_B200_FUSED": "1", "RSR_B200_PAIR": "12"}, MID),
    "f_pair_prof": ("librsr_b200_prof.so", {"RSR_B200_FUSED": "1", "RSR_B200_PAIR": "12"}, MID),
    "fp_t1n9": ("fp_t1n9.so", {"RSR_B200_FUSED": "1", "RSR_B200_PAIR": "12"}, MID),
    "fp_nr5": ("fp_nr5.so", {"RSR_B200_FUSED": "1", "RSR_B200_PAIR": "12"}, MID),
}
COMBINABLE = ("c0", "c3", "pk")   # build-flag variants (the f_* entries are launch plans of the same library)


def left():
    return BUDGET - (time.time() - T0)


def save():
    summary["elapsed_s"] = round(time.time() - T0, 1)
    with open(os.path.join(OUT, "final_round2_summary.json"), "w") as f:
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


def pytest_gpu(tag, lib=None):
    rc, so = run(f"pytest -m gpu ({tag})", [sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-q"], 150,
                 {"RSR_B200_LIB": lib} if lib else {}, f"pytest_gpu_{tag}.log")
    print("   ", (so.strip().splitlines() or ["-"])[-1], flush=True)
    return rc


def bench(tag, lib=None):
    rc, so = run(f"bench {tag}", [sys.executable, "bench.py", "--steps", "20", "--warmup", "3"], 90,
                 {"RSR_B200_LIB": lib} if lib else {}, f"bench_{tag}.json")
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
    lib, env, _ = ENTRIES[name]
    path = os.path.join(ABL, lib)
    if not os.path.exists(path) and not DRY:
        return None
    if DRY:
        import random
        us = {"pack_input": 52.0, "conv0": 83.0, "conv1|2|3": 560.0, "conv5|6|7": 300.0, "conv8|9": 270.0, "conv10+expand+skip": 213.0}
        if name == "pk_v2":
            us["pack_input"] = 40.0
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
    # 1. gates: the whole GPU suite on the shipped library, and on the previous default kept as fallback
    summary["pytest_shipped_rc"] = pytest_gpu("shipped")
    summary["pytest_fallback_rc"] = pytest_gpu("fallback_c0_r4n12", os.path.join(ABL, "c0_r4n12.so"))
    base1 = bench("shipped")
    # 2. hardware probes (tests/cuda): A operand from tensor memory; reference figures from the same box
    for probe in ("umma_ts_bench", "umma_bench"):
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
        for n, (lib, env, keys) in ENTRIES.items():
            if keys is None or not times[n]:
                continue
            v = sum(sum(t[k] for k in keys) for t in times[n]) / len(times[n])
            b = sum(bmean[k] for k in keys)
            gain = 1.0 - v / b
            ab[n] = {"launches": "+".join(keys), "us": round(v, 1), "base_us": round(b, 1), "gain": round(gain, 4), "runs": len(times[n])}
            print(f"    {n:12s} {'+'.join(keys):20s} {v:7.1f} us vs {b:7.1f}  gain {100 * gain:+.1f} %  ({len(times[n])} runs)", flush=True)
            if n[:2] in COMBINABLE and gain > 0.02 and (n[:2] not in winners or gain > winners[n[:2]][1]):
                winners[n[:2]] = (n, gain)
    summary["ab"] = ab
    summary["winners"] = {k: v[0] for k, v in winners.items()}
    save()
    # 4. winners combined into one library: GPU suite + bench against the shipped one
    if winners and left() > 100:
        flags = []
        for k, (n, _) in winners.items():
            flags += open(os.path.join(ABL, n + ".flags")).read().split()
        combo = os.path.join(ABL, "combo.so")
        nv = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", *flags, "-o", combo, "rusty_sr_b200/csrc/rsr_b200.cu", "-lcudart"]
        summary["combo_flags"] = flags
        rc, _ = run("build combo", nv, 120)
        if rc == 0 or DRY:
            summary["combo_tests_rc"] = pytest_gpu("combo", combo)
            if summary["combo_tests_rc"] == 0:
                c1 = bench("combo", combo)
                base2 = bench("shipped2") if left() > 70 else None
                c2 = bench("combo2", combo) if left() > 70 else None
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
    run("ncu full conv0", [ncu, "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:conv_strip", "-s", "4",
                           "-c", "1", "-o", os.path.join(OUT, "prof_conv0"), "-f", py, "tools/profile_step.py", "3"], 90, out="ncu_full_conv0.log")
    save()
    print(json.dumps(summary), flush=True)


if __name__ == "__main__":
    main()
