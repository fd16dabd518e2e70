#!/bin/bash
# Quick GPU iteration: parity tests, wait-cycle profile, bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
RSR_B200_LIB=$PWD/build/abl/librsr_b200_prof.so timeout 120 python tools/wait_profile.py > gpurun_out/waitprof.log 2>&1; cat gpurun_out/waitprof.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("ms/step", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roofline frac", round(d["roofline"]["frac"],3))
print({k:round(v*1e3,1) for k,v in d["kernel_ms"].items()})
print(d["clocks"], d.get("cpu_baseline"))
PY
tail -3 gpurun_out/bench.err
ABL_VARIANTS="8" bash tools/ablate.sh run
