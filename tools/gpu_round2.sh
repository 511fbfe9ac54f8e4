#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_widen.py -m gpu -q --durations=5 -p no:cacheprovider > gpurun_out/pytest_gpu_widen2.log 2>&1
echo "pytest widen exit $?" >> gpurun_out/pytest_gpu_widen2.log
tail -30 gpurun_out/pytest_gpu_widen2.log
timeout 100 python tools/widen_timing.py > gpurun_out/widen_timing2.log 2>&1; cat gpurun_out/widen_timing2.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"photometric|adam_multi|sh_view_split|stg_" -c 8 -o gpurun_out/prof_r1_widen python tools/widen_timing.py > gpurun_out/ncu_widen.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
