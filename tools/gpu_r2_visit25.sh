#!/bin/bash
# 2DGS family against the reference CUDA kernels, then the whole suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v25
timeout 400 python -m pytest tests/test_gpu_2dgs.py tests/test_gpu_png.py -q -s -p no:cacheprovider > ${P}_2dgs_tests.log 2>&1; echo "exit $?" >> ${P}_2dgs_tests.log; tail -60 ${P}_2dgs_tests.log
