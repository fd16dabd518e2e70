#!/bin/bash
# One GPU session: parity subset, then interleaved same-box A/B of library builds (per-launch times and whole-frame time).
# usage (under gpurun): tools/session_ab.sh TAG "pytest args" name=lib.so ...
cd "$(dirname "$0")/.."
TAG=$1; shift
PYT=$1; shift
mkdir -p gpurun_out
if [ -n "$PYT" ]; then
  timeout 900 python -m pytest $PYT -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
  cat gpurun_out/${TAG}_pytest.log
fi
timeout 600 python tools/ab_quick.py 3 "$@" 2>&1 | tee gpurun_out/${TAG}_ab.txt
for v in "$@"; do
  name=${v%%=*}; lib=${v#*=}
  if [ "$lib" = "shipped" ]; then lib=rusty_sr_b200/lib/librsr_b200.so; fi
  for i in 1 2; do
    RSR_B200_LIB=$PWD/$lib timeout 120 python tools/ab_frame.py 40 2>&1 | tail -1 | sed "s/^/$name  /" | tee -a gpurun_out/${TAG}_ab.txt
  done
done
