#!/bin/bash
# pipelined blend backward: blend tests + headline parity, stage timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v28
timeout 500 python -m pytest tests/test_gpu_blend_stream.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > ${P}_pytest.log 2>&1; echo "exit $?" >> ${P}_pytest.log; tail -4 ${P}_pytest.log
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; grep -v "^ref_\|^ref " ${P}_timing.log | tail -8
