#!/bin/bash
# moment-form blend backward: full GPU suite + stage timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v23
timeout 500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -5 ${P}_pytest_all.log
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; tail -14 ${P}_timing.log
