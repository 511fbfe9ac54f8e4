#!/bin/bash
# new single-process tests of the exchange kernels + whole suite + bench with all legs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v14
timeout 300 python -m pytest tests/test_gpu_peer.py -q -x -p no:cacheprovider > ${P}_peer.log 2>&1; echo "exit $?" >> ${P}_peer.log; tail -15 ${P}_peer.log | cut -c1-220
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -3 ${P}_pytest_all.log | cut -c1-200; grep -E "^FAILED|^ERROR" ${P}_pytest_all.log | head
timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err; tail -1 ${P}_bench.json | cut -c1-3000
