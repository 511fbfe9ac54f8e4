#!/bin/bash
# null record for odd trips, PTX predicate chains for the alpha tests, gate by equality, one vote per pair, -|T| as an OR
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v37
timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -4 ${P}_pytest_all.log
G='blend fwd max|blend bwd v_|^blend_|^isect|^proj|ours total'
run() { # name lib bwd_ctas fwd_ctas
  if [ -n "$2" ]; then export GSB200_LIB=$PWD/$2; else unset GSB200_LIB; fi
  if [ -n "$3" ]; then export GSB200_SBWD_CTAS=$3; else unset GSB200_SBWD_CTAS; fi
  if [ -n "$4" ]; then export GSB200_SFWD_CTAS=$4; else unset GSB200_SFWD_CTAS; fi
  timeout 150 python tools/gpu_stage_timing.py 3 > ${P}_timing_$1.log 2>&1
  echo "== $1: $(grep -E "$G" ${P}_timing_$1.log | tr '\n' ' ')"
}
run default "" "" ""
