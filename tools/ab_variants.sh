#!/bin/bash
# A/B timing of library variants built into gscodec_studio_b200/variants (see _build.py GSB200_LIB_OUT)
cd "$(dirname "$0")/.."
echo "== default"; python tools/gpu_stage_timing.py 3 2>&1 | grep -E "blend fwd max|blend bwd v_means2d|^blend_|^isect|ours total"
for lib in gscodec_studio_b200/variants/*.so; do
  echo "== $lib"
  GSB200_LIB=$PWD/$lib python tools/gpu_stage_timing.py 3 2>&1 | grep -E "blend fwd max|blend bwd v_means2d|^blend_|^isect|ours total"
done
