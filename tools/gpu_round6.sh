#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_all6.log 2>&1
echo "pytest all exit $?" >> gpurun_out/pytest_gpu_all6.log
tail -6 gpurun_out/pytest_gpu_all6.log
timeout 60 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench6.json
