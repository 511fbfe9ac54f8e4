#!/bin/bash
# backward CTA order by records walked: blend tests, timing A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v32
timeout 500 python -m pytest tests/test_gpu_blend_stream.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > ${P}_pytest.log 2>&1; echo "exit $?" >> ${P}_pytest.log; tail -3 ${P}_pytest.log
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; grep "^blend\|^isect" ${P}_timing.log
GSB200_BWD_ORDER=isects timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing_isects.log 2>&1; grep "^blend" ${P}_timing_isects.log
