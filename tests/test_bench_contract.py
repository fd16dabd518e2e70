"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints one JSON line with
the keys the driver reads, runs the CPU oracle only, and non-zero ranks stay silent."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600, env=e)


def test_reference_arm_json_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", env={"OMP_NUM_THREADS": "8"})
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["scaling"] == "weak"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "1920x1080" in d["config"]["workload"] and d["gpu_launches"] == 0


def test_reference_arm_only_rank0_works():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = run_bench("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CUDA device" in r.stdout
