#!/bin/bash
# vectorised fake-quant: tests; final launch list of a bench step; full bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v18
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -3 ${P}_pytest_all.log | cut -c1-200; grep -E "^FAILED|^ERROR" ${P}_pytest_all.log | head
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${P}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > ${P}_ncu_bench.log 2>&1; python tools/launch_breakdown.py ${P}_bench_launches.csv | head -40
timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err; tail -1 ${P}_bench.json | cut -c1-2500
