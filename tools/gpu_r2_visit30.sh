#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TILE_STATS=1 timeout 200 python tools/gpu_stage_timing.py 3 > gpurun_out/r2v30_tile_stats.log 2>&1; grep -v "^ref" gpurun_out/r2v30_tile_stats.log | tail -16
