#!/bin/bash
# round-2 visit 2: the stream blend -- its own tests first, a sanitizer pass, the whole GPU suite, A/B timing, one ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v2
timeout 300 python -m pytest tests/test_gpu_blend_stream.py -q -x -p no:cacheprovider > ${P}_stream_tests.log 2>&1; echo "exit $?" >> ${P}_stream_tests.log; tail -25 ${P}_stream_tests.log
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_blend_stream.py -q -x -p no:cacheprovider -k "long_lists or masks" > ${P}_memcheck.log 2>&1; echo "exit $?" >> ${P}_memcheck.log; tail -8 ${P}_memcheck.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -40 ${P}_pytest_all.log
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing_stream.log 2>&1; tail -22 ${P}_timing_stream.log
GSB200_BLEND=tile timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing_tile.log 2>&1; grep -E "^blend|ours total" ${P}_timing_tile.log
ONCE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:blend_stream -c 2 -o ${P}_blend_stream python tools/gpu_stage_timing.py 3 > ${P}_ncu.log 2>&1; echo "ncu exit $?"
