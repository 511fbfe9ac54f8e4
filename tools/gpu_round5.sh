#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_widen.py -m gpu -q --durations=6 -p no:cacheprovider > gpurun_out/pytest_gpu_widen5.log 2>&1
echo "pytest widen exit $?" >> gpurun_out/pytest_gpu_widen5.log
tail -40 gpurun_out/pytest_gpu_widen5.log
