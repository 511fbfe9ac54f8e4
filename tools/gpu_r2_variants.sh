#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== default"; python tools/gpu_stage_timing.py 3 2>&1 | grep -E "^blend_|ours total"
for lib in gscodec_studio_b200/variants/*.so; do
  echo "== $lib"
  GSB200_LIB=$PWD/$lib python tools/gpu_stage_timing.py 3 2>&1 | grep -E "^blend_|max abs diff"
done
