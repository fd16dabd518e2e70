#!/bin/bash
# One GPU session that regenerates the evidence the docs quote for the shipped library (run under gpurun, one GPU):
#   TAG_pytest.log             the whole `-m gpu` suite
#   TAG_bench.json             python bench.py (default flags) ; TAG_bench_reference.json  --impl reference
#   TAG_launches.csv           ncu launch list (gpu__time_duration.sum) of a short bench-like run (tools/profile_step.py)
#   TAG_prof_frame.ncu-rep     ncu --set full of the launches of one 1080p frame (read here with tools/ncu_summary.py)
#   TAG_sanitizer_<tool>.log   compute-sanitizer passes (tools/sanitize.sh)
# usage: tools/gpu_evidence.sh TAG [steps...]   steps default: pytest bench ncu sanitize
cd "$(dirname "$0")/.."
TAG=${1:-r03}; shift
STEPS=${*:-pytest bench ncu sanitize}
mkdir -p gpurun_out
for s in $STEPS; do
  case $s in
    pytest) timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log ;;
    bench)  timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench.err
            timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
            python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_bench_reference.json"):
    for l in open(f):
        try: d = json.loads(l)
        except Exception: continue
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "impl")}, "e2e", (d.get("e2e") or {}).get("value"), "launches", d.get("gpu_launches"))
PY
            ;;
    ncu)    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv \
                python tools/profile_step.py 6 > gpurun_out/${TAG}_ncu_launches.log 2>&1
            PROFILE_TIMING=0 timeout 900 ncu --set full --clock-control none --import-source on -s 8 -c 4 -f -o gpurun_out/${TAG}_prof_frame \
                python tools/profile_step.py 4 > gpurun_out/${TAG}_ncu_full.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_full.log ;;
    sanitize) SAN_TAG=${TAG} tools/sanitize.sh ;;
  esac
done
