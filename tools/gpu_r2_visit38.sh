#!/bin/bash
# which detail of the backward pair walk made visit 37 slower than visit 36 (554 vs 524 us)?  one switch at a time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v38
timeout 300 python -m pytest tests/test_gpu_blend_stream.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > ${P}_pytest.log 2>&1; echo "exit $?" >> ${P}_pytest.log; tail -3 ${P}_pytest.log
G='blend fwd max|blend bwd v_means2d|^blend_|ours total'
run() { # name lib
  if [ -n "$2" ]; then export GSB200_LIB=$PWD/$2; else unset GSB200_LIB; fi
  timeout 150 python tools/gpu_stage_timing.py 3 > ${P}_timing_$1.log 2>&1
  echo "== $1: $(grep -E "$G" ${P}_timing_$1.log | tr '\n' ' ')"
}
run default ""
for lib in gscodec_studio_b200/variants/*.so; do run $(basename $lib .so) $lib; done
