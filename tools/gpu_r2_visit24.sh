#!/bin/bash
# quantise on load: new tests first, then the full suite, then the bench both ways
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v24
timeout 300 python -m pytest tests/test_gpu_quant_on_load.py -q -x -p no:cacheprovider > ${P}_qol_tests.log 2>&1; echo "exit $?" >> ${P}_qol_tests.log; tail -30 ${P}_qol_tests.log
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_quant_on_load.py > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -5 ${P}_pytest_all.log
for flag in "" "--no-quant-on-load"; do
timeout 300 python bench.py --no-cpu-baseline --no-extra-legs $flag > ${P}_bench.json 2> ${P}_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v24_bench.json') if l.startswith('{')][-1])
print('N=1', d['steps'], d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'], d['gpu_launches'])
print({k:v for k,v in d['stage_ms'].items()})
PY
done
