#!/bin/bash
# two-records-per-trip blend walks (GSB_FWD_PAIR / GSB_BWD_PAIR): GPU suite on the default build, then A/B stage timing
# of the variants in gscodec_studio_b200/variants (built with GSB200_LIB_OUT / GSB200_NVCC_EXTRA, see _build.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v34
timeout 400 python -m pytest tests -m gpu -q -x -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -4 ${P}_pytest_all.log
G='blend fwd max|blend bwd v_|^blend_|ours total'
run() { # name lib ctas
  if [ -n "$2" ]; then export GSB200_LIB=$PWD/$2; else unset GSB200_LIB; fi
  if [ -n "$3" ]; then export GSB200_SBWD_CTAS=$3; else unset GSB200_SBWD_CTAS; fi
  timeout 150 python tools/gpu_stage_timing.py 3 > ${P}_timing_$1.log 2>&1
  echo "== $1: $(grep -E "$G" ${P}_timing_$1.log | tr '\n' ' ')"
}
run pair_default "" ""
run a_nopair gscodec_studio_b200/variants/a_nopair.so ""
run d_pair_b6_c6 gscodec_studio_b200/variants/d_pair_b6.so 6
run e_pair_b7_c5 gscodec_studio_b200/variants/e_pair_b7.so 5
run e_pair_b7_c6 gscodec_studio_b200/variants/e_pair_b7.so 6
run e_pair_b7_c8 gscodec_studio_b200/variants/e_pair_b7.so 8
