#!/bin/bash
# fused replica scan: full suite, stage timing (default / legacy scan / two pixel pairs per thread)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v27
timeout 500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -4 ${P}_pytest_all.log
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; grep -v "^ref_\|^ref " ${P}_timing.log | tail -8
GSB200_ISECT_SCAN=legacy timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing_legacy.log 2>&1; grep "^isect" ${P}_timing_legacy.log
GSB200_STREAM_NP=2 timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing_np2.log 2>&1; grep "^blend" ${P}_timing_np2.log
