#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v7
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -8 ${P}_pytest_all.log | cut -c1-300; grep -E "^FAILED|^ERROR" ${P}_pytest_all.log | head
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ${P}_bench.json 2> ${P}_bench.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v7_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['api_literal']['value'], d['entropy_on']['value'], d['ref_cuda'].get('value'), d['roofline']['frac'])
print(d['stage_ms'])
PY
