#!/bin/bash
# compute-sanitizer memcheck over the two-record blend kernels and the private-histogram isect kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v48
timeout 85 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_blend_stream.py -q -x -p no:cacheprovider -k "long_lists or masks" > ${P}_memcheck_blend.log 2>&1; echo "blend memcheck exit $?"; grep -E "passed|failed|ERROR SUMMARY" ${P}_memcheck_blend.log | tail -3
timeout 55 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "private_histograms" > ${P}_memcheck_isect.log 2>&1; echo "isect memcheck exit $?"; grep -E "passed|failed|ERROR SUMMARY" ${P}_memcheck_isect.log | tail -3
