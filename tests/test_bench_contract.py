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


def test_per_launch_roofline_table():
    """The per-launch roofline of the bench line, on the launch times of a recorded run (profiles/r03_bench.json)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    rec = json.load(open(os.path.join(ROOT, "profiles", "r03_bench.json")))
    peaks = {"bf16_tflops": 1361.8, "bf16_tflops_burst": 1621.1, "hbm_gbs": 6550.4}
    t = b.per_launch_roofline(rec["kernel_ms"], 1920 * 1080, peaks, mma_per_chunk=2)
    assert set(t) == set(rec["kernel_ms"]) == {"conv0", "conv1|2|3", "conv5|6|7+conv8|9 (fused)", "conv10+expand+skip"}
    # whole-net useful FLOPs and bytes of the launch table agree with SURVEY.md 8(a) / DESIGN.md (fused plan: 768 B/px less)
    fused = [k for k in b.LAUNCH_WORK if k not in ("conv5|6|7", "conv8|9")]
    assert sum(b.LAUNCH_WORK[k][0] for k in fused) == b.FLOPS_PER_LR_PX
    assert sum(b.LAUNCH_WORK[k][1] for k in b.LAUNCH_WORK if "fused" not in k) == 2456
    assert sum(b.LAUNCH_WORK[k][1] for k in fused) == 1688
    assert t["conv1|2|3"]["bound"] == "tensor" and 0.5 < t["conv1|2|3"]["frac"] < 0.6 and 0.9 < t["conv1|2|3"]["frac_executed"] <= 1.0
    assert t["conv0"]["bound"] == "hbm" and all(0 < v["frac"] <= 1.0 for v in t.values())
    assert b.per_launch_roofline({"conv0": float("nan"), "unknown": 1.0}, 10, peaks) == {}


def test_recorded_bench_line_has_the_contract_keys_and_matches_the_docs():
    """profiles/r03_bench.json (the recorded line of the shipped default plan) carries every key the driver and the judge read,
    and the headline numbers quoted in README.md / DESIGN.md are the ones in that file."""
    rec = json.load(open(os.path.join(ROOT, "profiles", "r03_bench.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in rec, k
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(rec["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(rec["cpu_baseline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(rec["e2e"])
    assert rec["gpu_launches"] == 4 * rec["steps"] and rec["config"]["workload"].startswith("1920x1080")
    assert abs(rec["value"] - 18.6624 / rec["ms_per_step"] * 1e3) < 1.0          # Mpix/s = output Mpix per frame / s per frame
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(rec["clocks"]["reasons"])
    ms = f"{rec['ms_per_step']:.3f}"
    for doc in ("DESIGN.md", os.path.join("profiles", "README.md")):
        assert ms in open(os.path.join(ROOT, doc)).read(), (doc, ms)
