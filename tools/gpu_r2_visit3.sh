#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v3
timeout 300 python -m pytest tests/test_gpu_blend_stream.py -q -x -p no:cacheprovider > ${P}_stream_tests.log 2>&1; echo "exit $?" >> ${P}_stream_tests.log; tail -5 ${P}_stream_tests.log
GSB200_STREAM_NP=2 timeout 300 python -m pytest tests/test_gpu_blend_stream.py -q -x -p no:cacheprovider > ${P}_stream_tests_np2.log 2>&1; echo "exit $?" >> ${P}_stream_tests_np2.log; tail -5 ${P}_stream_tests_np2.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_widen.py -q -p no:cacheprovider -k "headline or hashgrid or gaussian_model or simulation_matches_reference or golden" > ${P}_fixed_tests.log 2>&1; echo "exit $?" >> ${P}_fixed_tests.log; tail -12 ${P}_fixed_tests.log
for np in 1 2; do
  GSB200_STREAM_NP=$np timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing_np$np.log 2>&1; echo "== NP=$np"; grep -E "^blend|max abs|rel err|ours total" ${P}_timing_np$np.log
done
