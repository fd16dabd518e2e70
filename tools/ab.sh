#!/bin/bash
# A/B timing of library variants in build/abl/*.so (interleaved, 3 rounds, same box); run under gpurun.
cd "$(dirname "$0")/.."
for round in 1 2 3; do
  for so in build/abl/*.so; do
    RSR_B200_LIB=$PWD/$so python tools/profile_step.py 8 2>&1 | tail -1 | python -c "
import sys, ast
d = ast.literal_eval(sys.stdin.read())
print('%-16s' % '$(basename $so .so)', ' '.join('%s %.1f' % (k.split('+')[0][:9], v*1e3) for k, v in d.items()), ' total %.1f' % (sum(d.values())*1e3))"
  done
done
