#!/bin/bash
# Builds timing-ablation variants of the library (wrong results on purpose) plus the wait-cycle
# profiling variant.  `tools/ablate.sh build` here (no GPU needed), `tools/ablate.sh run` under gpurun.
cd "$(dirname "$0")/.."
VARIANTS="${ABL_VARIANTS:-0 8 16 27}"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared"
if [ "$1" = "build" ]; then
  mkdir -p build/abl
  for v in $VARIANTS; do
    $NV -DRSR_ABL=$v -o build/abl/librsr_b200_abl$v.so rusty_sr_b200/csrc/rsr_b200.cu -lcudart &
  done
  $NV -DRSR_PROFILE=1 -DRSR_FUSED_PROF=1 -o build/abl/librsr_b200_prof.so rusty_sr_b200/csrc/rsr_b200.cu -lcudart &
  wait; ls build/abl
else
  for v in $VARIANTS; do
    RSR_B200_LIB=$PWD/build/abl/librsr_b200_abl$v.so python tools/profile_step.py 4 2>&1 | tail -1 | sed "s/^/ABL=$v  /"
  done
fi
