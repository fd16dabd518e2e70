#!/bin/bash
# Third GPU session of round 1 (short): is the fused middle launch on CTA pairs faster than the split plan under
# bench.py's sustained load?  Interleaved bench runs, the whole GPU suite with that plan as the default, per-launch
# A/B lines, one full ncu capture of the fused pair kernel.
mkdir -p gpurun_out
B() { # tag, env...
  tag=$1; shift
  env "$@" timeout 60 python bench.py --steps 20 --warmup 3 > gpurun_out/bench3_$tag.json 2> gpurun_out/bench3_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench3_{sys.argv[1]}.json"))
    print(f"bench {sys.argv[1]:10s} ms/step {d['ms_per_step']:.4f} value {d['value']:.1f} e2e {d['e2e']['value']:.1f}",
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()}, d["clocks"]["reasons"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("bench parse failed", sys.argv[1], e)
PY
}
B split1 RSR_B200_FUSED=0
B fpair1 RSR_B200_FUSED=1 RSR_B200_PAIR=12
B split2 RSR_B200_FUSED=0
B fpair2 RSR_B200_FUSED=1 RSR_B200_PAIR=12
RSR_B200_FUSED=1 RSR_B200_PAIR=12 timeout 100 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_fpair_default.log 2>&1; echo "pytest (fused pair as default) rc=$?"; tail -2 gpurun_out/pytest_gpu_fpair_default.log
for i in 1 2; do
  timeout 30 python tools/ab_step.py 6 | tee -a gpurun_out/ab3.jsonl | cut -c1-300
  RSR_B200_FUSED=1 RSR_B200_PAIR=12 timeout 30 python tools/ab_step.py 6 | tee -a gpurun_out/ab3.jsonl | cut -c1-300
done
RSR_B200_FUSED=1 RSR_B200_PAIR=12 timeout 60 ncu --set full --clock-control none --import-source on -k regex:conv_fused -s 1 -c 1 -o gpurun_out/prof_fused_pair -f python tools/profile_step.py 3 > gpurun_out/ncu_fused_pair.log 2>&1; echo "ncu rc=$?"
