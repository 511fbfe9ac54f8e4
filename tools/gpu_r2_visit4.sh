#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v4
timeout 300 python -m pytest tests/test_gpu_blend_stream.py -q -x -p no:cacheprovider > ${P}_stream_tests.log 2>&1; echo "exit $?" >> ${P}_stream_tests.log; tail -5 ${P}_stream_tests.log
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; grep -E "^blend|max abs|rel err|ours total" ${P}_timing.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_widen.py -q -p no:cacheprovider -k "headline or hashgrid_encoder_matches_reference" > ${P}_fixed_tests.log 2>&1; echo "exit $?" >> ${P}_fixed_tests.log; grep -E "^E  |passed|failed|headline from" ${P}_fixed_tests.log | head -20
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > ${P}_bench.json 2> ${P}_bench.err; echo "bench exit $?"; cat ${P}_bench.json; tail -5 ${P}_bench.err
