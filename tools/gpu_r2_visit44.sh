#!/bin/bash
# tile intersection with CTA-private shared-memory histograms (one global atomic per (CTA, tile) pair instead of one per
# intersection in the count and scatter passes): bit-exact suite, stage timing A/B, bench A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v44
timeout 500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -4 ${P}_pytest_all.log
G='radii equal|^isect|^blend_|ours total'
for m in 0 1; do
  GSB200_ISECT_PRIVATE=$m timeout 150 python tools/gpu_stage_timing.py 3 > ${P}_timing_priv$m.log 2>&1
  echo "== private $m: $(grep -E "$G" ${P}_timing_priv$m.log | tr '\n' ' ')"
done
for m in 0 1; do
  GSB200_ISECT_PRIVATE=$m timeout 300 python bench.py --no-cpu-baseline --no-extra-legs > ${P}_bench_priv$m.json 2> ${P}_bench_priv$m.err; echo "private $m: bench exit $?"
  python - $m <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/r2v44_bench_priv%s.json'%sys.argv[1]) if l.startswith('{')][-1])
print('  value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], {k:v for k,v in d['stage_ms'].items() if k.startswith('isect')})
PY
done
