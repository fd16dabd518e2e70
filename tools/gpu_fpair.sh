#!/bin/bash
# GPU check of the pair-fused middle launch (RSR_B200_PAIR bit 3 + fused plan): tests, then same-box A/B.
mkdir -p gpurun_out
RSR_B200_PAIR=12 timeout 120 python tools/fused_debug.py > gpurun_out/fpair_debug.log 2>&1; echo "debug(single-CTA dbg kernel) rc=$?"; grep -E "OUTPUT|EXCEPTION" gpurun_out/fpair_debug.log
RSR_B200_PAIR=12 timeout 300 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/pytest_fpair.log 2>&1; echo "pytest fused on pairs rc=$?"; tail -12 gpurun_out/pytest_fpair.log
for cfg in "4 split" "4 fused" "12 fused" "4 split" "12 fused"; do
  set -- $cfg
  RSR_B200_PAIR=$1 timeout 200 python bench.py --steps 20 --warmup 3 --plan $2 > gpurun_out/bench_p$1_$2.json 2> gpurun_out/bench_p$1_$2.err; echo "bench pair=$1 plan=$2 rc=$?"
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_p{sys.argv[1]}_{sys.argv[2]}.json"))
    print("  ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()}, d["clocks"]["reasons"])
except Exception as e:
    print("bench parse failed", e)
PY
  tail -2 gpurun_out/bench_p$1_$2.err
done
