#!/bin/bash
# Quick GPU check of the fused middle launch: diagnostic, fused tests, same-box A/B bench.
mkdir -p gpurun_out
timeout 200 python tools/fused_debug.py --profile > gpurun_out/fused_debug.log 2>&1; echo "debug rc=$?"; grep -E "max\|err\||EXCEPTION|OUTPUT|%|total|kernel ms" gpurun_out/fused_debug.log | head -40
timeout 400 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/pytest_fused.log 2>&1; echo "pytest fused rc=$?"; tail -8 gpurun_out/pytest_fused.log
for plan in split fused; do
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
