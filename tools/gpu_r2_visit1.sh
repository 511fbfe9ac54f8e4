#!/bin/bash
# round-2 visit 1: pin the hash-grid oracle, run the gated hash-grid + radix-sort tests, keep a stage-timing log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2v1_gpu.txt
timeout 200 python -m oracle.hashgrid > gpurun_out/r2v1_hashgrid_pin.log 2>&1; echo "pin exit $?" >> gpurun_out/r2v1_hashgrid_pin.log
tail -8 gpurun_out/r2v1_hashgrid_pin.log
GSB200_EXPERIMENTAL_HASHGRID=1 timeout 300 python -m pytest tests/test_gpu_widen.py -k "hashgrid or gaussian_model" -q -p no:cacheprovider > gpurun_out/r2v1_hashgrid_test.log 2>&1
echo "exit $?" >> gpurun_out/r2v1_hashgrid_test.log; tail -30 gpurun_out/r2v1_hashgrid_test.log
GSB200_TILE_SORT=radix timeout 300 python -m pytest tests/test_gpu_parity.py -k "isect" -q -p no:cacheprovider > gpurun_out/r2v1_radix_test.log 2>&1
echo "exit $?" >> gpurun_out/r2v1_radix_test.log; tail -15 gpurun_out/r2v1_radix_test.log
timeout 200 python tools/gpu_stage_timing.py 3 > gpurun_out/r2v1_stage_timing.log 2>&1; tail -30 gpurun_out/r2v1_stage_timing.log
GSB200_TILE_SORT=radix timeout 200 python tools/gpu_stage_timing.py 3 > gpurun_out/r2v1_stage_timing_radix.log 2>&1; grep -E "^isect|ours total" gpurun_out/r2v1_stage_timing_radix.log
