#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --durations=6 -p no:cacheprovider -x > gpurun_out/pytest_gpu_all4.log 2>&1
echo "pytest all exit $?" >> gpurun_out/pytest_gpu_all4.log
tail -22 gpurun_out/pytest_gpu_all4.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke4.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke4.log; tail -2 gpurun_out/smoke4.log
timeout 200 python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench exit $?"; cat gpurun_out/bench4.json; tail -3 gpurun_out/bench4.err
timeout 100 python tools/widen_timing.py > gpurun_out/widen_timing4.log 2>&1; cat gpurun_out/widen_timing4.log
