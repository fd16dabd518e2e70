#!/bin/bash
# Builds library variants for same-box A/B timing into build/abl/ (no GPU needed; run tools/ab.sh under gpurun).
# usage: tools/ab_build.sh name1 "<nvcc -D flags>" name2 "<flags>" ...   ("base" with "" = the shipped configuration)
cd "$(dirname "$0")/.."
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared"
mkdir -p build/abl
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  echo "$flags" > build/abl/$name.flags
  ( $NV $flags -o build/abl/$name.so rusty_sr_b200/csrc/rsr_b200.cu -lcudart 2> build/abl/$name.log || echo "BUILD FAILED: $name (see build/abl/$name.log)" ) &
done
wait
ls -la build/abl/*.so
