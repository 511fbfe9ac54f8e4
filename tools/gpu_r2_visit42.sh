#!/bin/bash
# colour stage on a side stream (forward beside the tile intersection, backward beside the projection backward): suite + A/B bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v42
timeout 500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -4 ${P}_pytest_all.log
for m in 0 1 0 1; do
  GSB200_COLOR_STREAM=$m timeout 300 python bench.py --no-cpu-baseline --no-extra-legs > ${P}_bench_cs$m.json 2> ${P}_bench_cs$m.err; echo "colour stream $m: bench exit $?"
  python - $m <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/r2v42_bench_cs%s.json'%sys.argv[1]) if l.startswith('{')][-1])
print('  value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['launch_ms'], 'loss', d['loss'])
PY
done
