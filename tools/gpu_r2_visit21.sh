#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v21
for i in 1 2 3 4; do
timeout 600 python bench.py --no-cpu-baseline --no-extra-legs > ${P}_bench.json 2> ${P}_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v21_bench.json') if l.startswith('{')][-1])
print('N=1', d['steps'], d['value'], d['ms_per_step'], d['e2e']['value'])
PY
done
