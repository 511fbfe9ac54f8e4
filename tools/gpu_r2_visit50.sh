#!/bin/bash
# backward pair walk at 8 (62 registers) / 7 (69, the build of the bench records) / 6 (76) resident CTAs per SM by launch bounds; 8 ships
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v50
G='blend bwd v_means2d|^blend_bwd|ours total'
run() { # name lib
  if [ -n "$2" ]; then export GSB200_LIB=$PWD/$2; else unset GSB200_LIB; fi
  timeout 60 python tools/gpu_stage_timing.py 3 > ${P}_timing_$1.log 2>&1
  echo "== $1: $(grep -E "$G" ${P}_timing_$1.log | tr '\n' ' ')"
}
for lib in gscodec_studio_b200/variants/*.so; do run $(basename $lib .so) $lib; done
