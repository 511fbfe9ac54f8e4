#!/bin/bash
# colour stage inside the intersection-count wait + host profile of one step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v13
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -3 ${P}_pytest_all.log | cut -c1-200; grep -E "^FAILED|^ERROR" ${P}_pytest_all.log | head
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra-legs > ${P}_bench.json 2> ${P}_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v13_bench.json') if l.startswith('{')][-1])
print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
PY
timeout 300 python tools/cpu_overhead.py > ${P}_host.log 2>&1; grep -E "async|sync every" ${P}_host.log
