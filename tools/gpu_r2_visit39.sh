#!/bin/bash
# bisect of visit 37's slowdown on the v36 source: y1 = v36, y2 = + 48 bytes of shared memory, y4 = validity as !(a < b || c < d)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v39
G='blend fwd max|blend bwd v_means2d|^blend_|ours total'
run() { # name lib
  if [ -n "$2" ]; then export GSB200_LIB=$PWD/$2; else unset GSB200_LIB; fi
  timeout 150 python tools/gpu_stage_timing.py 3 > ${P}_timing_$1.log 2>&1
  echo "== $1: $(grep -E "$G" ${P}_timing_$1.log | tr '\n' ' ')"
}
for lib in gscodec_studio_b200/variants/*.so; do run $(basename $lib .so) $lib; done
