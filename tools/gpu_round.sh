#!/bin/bash
# One GPU session: bring-up diagnostic, parity tests, bench, ncu launch list + full capture.
mkdir -p gpurun_out
timeout 300 python tests/gpu_bringup.py > gpurun_out/bringup.log 2>&1; echo "bringup rc=$?"; grep -E "OUTPUT|BRINGUP|EXCEPTION|kernel ms" gpurun_out/bringup.log | head -20
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_strip -s 6 -c 5 -o gpurun_out/prof_conv -f python tools/profile_step.py 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
