#!/bin/bash
# compute-sanitizer memcheck: colour side stream, blend forward / backward variants (absgrad, backgrounds, channels) of the parity suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v49
timeout 60 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_colour_stream.py -q -x -p no:cacheprovider > ${P}_memcheck_colour.log 2>&1; echo "colour memcheck exit $?"; grep -E "passed|failed|ERROR SUMMARY" ${P}_memcheck_colour.log | tail -3
timeout 60 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_blend_stream.py -q -x -p no:cacheprovider -k "blend_fwd_bwd or absgrad or equals_tile_kernels" > ${P}_memcheck_blend2.log 2>&1; echo "blend memcheck exit $?"; grep -E "passed|failed|ERROR SUMMARY" ${P}_memcheck_blend2.log | tail -3
