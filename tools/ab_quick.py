"""Interleaved same-box A/B of library variants (run under gpurun): per-launch CUDA-event times from tools/ab_step.py.
usage: python tools/ab_quick.py ROUNDS name=lib.so[,ENV=VAL...] name2=lib2.so ...     (lib paths relative to the repo root;
"shipped" = rusty_sr_b200/lib/librsr_b200.so).  Prints one line per variant: mean microseconds per launch and total."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rounds = int(sys.argv[1])
variants = []
for a in sys.argv[2:]:
    name, rest = a.split("=", 1)
    parts = rest.split(",")
    lib = os.path.join(ROOT, "rusty_sr_b200", "lib", "librsr_b200.so") if parts[0] == "shipped" else os.path.join(ROOT, parts[0])
    env = dict(kv.split("=", 1) for kv in parts[1:])
    variants.append((name, lib, env))
acc = {n: [] for n, _, _ in variants}
for r in range(rounds):
    for name, lib, env in variants:
        e = dict(os.environ)
        e.update(env)
        e["RSR_B200_LIB"] = lib
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ab_step.py"), "6"], env=e, stdout=subprocess.PIPE,
                                 stderr=subprocess.PIPE, text=True, timeout=120)
            acc[name].append(json.loads(out.stdout.strip().splitlines()[-1])["us"])
        except Exception as ex:  # noqa
            print(f"{name}: run failed: {ex}", flush=True)
for name, runs in acc.items():
    if not runs:
        continue
    mean = {k: sum(r[k] for r in runs) / len(runs) for k in runs[0]}
    print(f"{name:14s} total {sum(mean.values()):7.1f} us | " + "  ".join(f"{k} {v:6.1f}" for k, v in mean.items()) + f"  ({len(runs)} runs)", flush=True)
