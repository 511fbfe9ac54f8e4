#!/bin/bash
# bucket-sort edge cases (+ memcheck on a subset), e2e with the split input wait
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v15
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "depth_bucket" > ${P}_bucket.log 2>&1; echo "exit $?" >> ${P}_bucket.log; tail -12 ${P}_bucket.log | cut -c1-220
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer.py -q -x -p no:cacheprovider -k "(depth_bucket and (1500 or 5000 or 14000)) or exchange or barrier" > ${P}_memcheck.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" ${P}_memcheck.log | tail -3
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "depth_bucket and (1500 or 5000) and (uniform or equal)" > ${P}_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed" ${P}_racecheck.log | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-extra-legs > ${P}_bench.json 2> ${P}_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v15_bench.json') if l.startswith('{')][-1])
print('N=1', d['steps'], d['value'], d['ms_per_step'], d['e2e']['value'])
PY
