#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_widen.py -m gpu -q --durations=5 -p no:cacheprovider > gpurun_out/pytest_gpu_widen3.log 2>&1
echo "pytest widen exit $?" >> gpurun_out/pytest_gpu_widen3.log
tail -30 gpurun_out/pytest_gpu_widen3.log
timeout 200 python bench.py > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit $?"; cat gpurun_out/bench3.json; tail -3 gpurun_out/bench3.err
timeout 100 python tools/widen_timing.py > gpurun_out/widen_timing3.log 2>&1; cat gpurun_out/widen_timing3.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"photometric" -s 6 -c 2 -o gpurun_out/prof_r1_loss python tools/widen_timing.py > gpurun_out/ncu_loss.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
