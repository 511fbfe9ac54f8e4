#!/bin/bash
# final ncu --set full of the two stream blend kernels (one launch each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v22
ONCE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:blend_stream -c 2 -o ${P}_blend_stream python tools/gpu_stage_timing.py 3 > ${P}_ncu.log 2>&1; ls -la ${P}_blend_stream.ncu-rep
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; grep -E "^[a-z_]+ +[0-9.]+ us|ours total|ref total" ${P}_timing.log | head -20
