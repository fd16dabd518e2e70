#!/bin/bash
# One GPU session for the fused middle launch: bring-up (split + fused), fused tests, same-box A/B bench of
# the two launch plans, the parity suite, ncu launch list + one full capture of the fused kernel.
mkdir -p gpurun_out
timeout 240 python tests/gpu_bringup.py > gpurun_out/bringup.log 2>&1; echo "bringup rc=$?"
grep -E "OUTPUT|BRINGUP|EXCEPTION|l3_activ|bad " gpurun_out/bringup.log | head -40
timeout 400 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/pytest_fused.log 2>&1; echo "pytest fused rc=$?"; tail -15 gpurun_out/pytest_fused.log
for plan in split fused split fused; do
  timeout 200 python bench.py --steps 20 --warmup 3 --plan $plan > gpurun_out/bench_$plan.json 2> gpurun_out/bench_$plan.err; echo "bench $plan rc=$?"
  python - "$plan" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()}, d["clocks"]["reasons"])
except Exception as e:
    print("bench parse failed", e)
PY
done
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest parity rc=$?"; tail -5 gpurun_out/pytest_parity.log
RSR_B200_FUSED=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_fused.csv python tools/profile_step.py 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
RSR_B200_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_fused -s 1 -c 1 -o gpurun_out/prof_fused -f python tools/profile_step.py 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
RSR_B200_FUSED=1 timeout 240 compute-sanitizer --tool memcheck --error-exitcode 7 python - > gpurun_out/sanitizer_fused_memcheck.log 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import rusty_sr_b200 as R
rng = np.random.default_rng(0)
img = rng.integers(0, 256, size=(37, 150, 3), dtype=np.uint8)
with R.Upscaler(parameters="anime") as up:
    a = up.forward_u8(img)
    print("ok", a.shape, up.last_launch_count())
PY
echo "memcheck fused rc=$?"; grep -E "ERROR SUMMARY|ok \(|Error" gpurun_out/sanitizer_fused_memcheck.log | head -5
