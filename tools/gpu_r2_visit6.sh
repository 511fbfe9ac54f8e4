#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v6
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -25 ${P}_pytest_all.log | cut -c1-300
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "packed or tile_sizes" > ${P}_memcheck.log 2>&1; echo "exit $?" >> ${P}_memcheck.log; grep -E "ERROR SUMMARY|passed|failed" ${P}_memcheck.log | tail -4
