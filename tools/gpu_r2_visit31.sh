#!/bin/bash
# segment-parallel blend backward: stream-vs-tile tests, parity suite, timing with and without segments
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v31
timeout 500 python -m pytest tests/test_gpu_blend_stream.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > ${P}_pytest.log 2>&1; echo "exit $?" >> ${P}_pytest.log; grep -n "rel err\|assert\|Error" ${P}_pytest.log | head -20; tail -4 ${P}_pytest.log
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; grep -v "^ref_\|^ref " ${P}_timing.log | tail -10
GSB200_BWD_SEGMENTS=0 timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing_noseg.log 2>&1; grep "^blend" ${P}_timing_noseg.log
