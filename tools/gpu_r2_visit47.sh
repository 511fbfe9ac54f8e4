#!/bin/bash
# forward with 7 / 8 resident CTAs per SM (56 / 48 registers, a few bytes of spill): the consumers of the two-record kernel
# spend 28 % of their samples waiting for the producer -- do more co-resident tiles fill those gaps?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v47
G='blend fwd max|^blend_fwd|ours total'
run() { # name lib
  if [ -n "$2" ]; then export GSB200_LIB=$PWD/$2; else unset GSB200_LIB; fi
  timeout 150 python tools/gpu_stage_timing.py 3 > ${P}_timing_$1.log 2>&1
  echo "== $1: $(grep -E "$G" ${P}_timing_$1.log | tr '\n' ' ')"
}
run default ""
for lib in gscodec_studio_b200/variants/*.so; do run $(basename $lib .so) $lib; done
