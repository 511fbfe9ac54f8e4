#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v9
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -4 ${P}_pytest_all.log | cut -c1-200; grep -E "^FAILED|^ERROR" ${P}_pytest_all.log | head
timeout 100 python -m pytest tests/test_gpu_parity.py -q -s -p no:cacheprovider -k headline 2>&1 | grep -E "headline from|grad |passed|failed" > ${P}_headline.log; cat ${P}_headline.log
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; grep -E "^blend|^isect|^proj|bit-equal|radii equal|rel err|ours total" ${P}_timing.log
