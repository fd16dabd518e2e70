#!/bin/bash
# GPU check of the CTA-pair kernels: bring-up per launch mask, then parity + A/B bench for the masks that pass.
mkdir -p gpurun_out
for mask in ${PAIR_MASKS:-4}; do
  RSR_B200_PAIR=$mask timeout 120 python tests/gpu_bringup.py > gpurun_out/bringup_pair$mask.log 2>&1; echo "bringup pair mask $mask rc=$?"
  grep -E "max\|err\||BRINGUP|EXCEPTION|bad " gpurun_out/bringup_pair$mask.log | grep -v fused | head -${BRINGUP_LINES:-14}
done
if [ -n "$BENCH_MASKS" ]; then
  for mask in $BENCH_MASKS; do
    RSR_B200_PAIR=$mask timeout 200 python bench.py --steps 20 --warmup 3 --plan split > gpurun_out/bench_pair$mask.json 2> gpurun_out/bench_pair$mask.err; echo "bench pair mask $mask rc=$?"
    python - "$mask" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_pair{sys.argv[1]}.json"))
    print("mask", sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()}, d["clocks"]["reasons"])
except Exception as e:
    print("bench parse failed", e)
PY
    tail -2 gpurun_out/bench_pair$mask.err
  done
fi
if [ -n "$TEST_MASK" ]; then
  RSR_B200_PAIR=$TEST_MASK timeout 500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_pair.log 2>&1; echo "pytest pair mask $TEST_MASK rc=$?"; tail -6 gpurun_out/pytest_pair.log
fi
