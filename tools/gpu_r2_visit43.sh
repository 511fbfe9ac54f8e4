#!/bin/bash
# forward ring: stage size 64 / 48 / 32 records (a consumer starts on a stage when the producer has completed it), consumer back-off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v43
GSB200_LIB=$PWD/gscodec_studio_b200/variants/s32.so timeout 300 python -m pytest tests/test_gpu_blend_stream.py tests/test_gpu_parity.py tests/test_gpu_colour_stream.py -m gpu -q -x -p no:cacheprovider > ${P}_pytest_s32.log 2>&1; echo "exit $?" >> ${P}_pytest_s32.log; tail -3 ${P}_pytest_s32.log
G='blend fwd max|blend bwd v_means2d|^blend_|ours total'
run() { # name lib
  if [ -n "$2" ]; then export GSB200_LIB=$PWD/$2; else unset GSB200_LIB; fi
  timeout 150 python tools/gpu_stage_timing.py 3 > ${P}_timing_$1.log 2>&1
  echo "== $1: $(grep -E "$G" ${P}_timing_$1.log | tr '\n' ' ')"
}
run default ""
for lib in gscodec_studio_b200/variants/*.so; do run $(basename $lib .so) $lib; done
