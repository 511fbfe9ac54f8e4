#!/bin/bash
# One GPU visit: parity tests, smoke, headline bench, widen timings, launch list.  Most important first; every
# step bounded by its own timeout so that a hang cannot eat the box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_widen.py -m gpu -q --durations=5 -p no:cacheprovider > gpurun_out/pytest_gpu_widen.log 2>&1
echo "pytest widen exit $?" >> gpurun_out/pytest_gpu_widen.log
tail -40 gpurun_out/pytest_gpu_widen.log
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest parity exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 100 python bench.py --cat-colors --no-cpu-baseline > gpurun_out/bench_cat.json 2> gpurun_out/bench_cat.err; cat gpurun_out/bench_cat.json
timeout 100 python tools/widen_timing.py > gpurun_out/widen_timing.log 2>&1; cat gpurun_out/widen_timing.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"
