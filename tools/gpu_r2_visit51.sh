#!/bin/bash
# blend tests on the 8-CTA backward build (last GPU seconds of the round)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 45 python -m pytest tests/test_gpu_blend_stream.py tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "stream or blend_fwd_bwd or absgrad or rasterization_matches" > gpurun_out/r2v51_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2v51_pytest.log; tail -3 gpurun_out/r2v51_pytest.log
