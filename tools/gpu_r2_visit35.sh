#!/bin/bash
# forward: 2 / 3 / 4 records per trip (GSB_FWD_GROUP), backward pair walk with unconditional publish; resident-CTA caps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v35
timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -4 ${P}_pytest_all.log
G='blend fwd max|blend bwd v_means2d|^blend_|ours total'
run() { # name lib bwd_ctas fwd_ctas
  if [ -n "$2" ]; then export GSB200_LIB=$PWD/$2; else unset GSB200_LIB; fi
  if [ -n "$3" ]; then export GSB200_SBWD_CTAS=$3; else unset GSB200_SBWD_CTAS; fi
  if [ -n "$4" ]; then export GSB200_SFWD_CTAS=$4; else unset GSB200_SFWD_CTAS; fi
  timeout 150 python tools/gpu_stage_timing.py 3 > ${P}_timing_$1.log 2>&1
  echo "== $1: $(grep -E "$G" ${P}_timing_$1.log | tr '\n' ' ')"
}
run default "" "" ""
run default_c7 "" 7 ""
run default_c6_f5 "" 6 5
run f3b5 gscodec_studio_b200/variants/f3b5.so "" ""
run f3b6 gscodec_studio_b200/variants/f3b6.so "" ""
run f4b5 gscodec_studio_b200/variants/f4b5.so "" ""
