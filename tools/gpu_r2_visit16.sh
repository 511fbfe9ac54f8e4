#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SCENE_GRID=1 timeout 300 python tools/cpu_overhead.py > gpurun_out/r2v16_host_small.log 2>&1; grep -E "async|sync every" gpurun_out/r2v16_host_small.log
