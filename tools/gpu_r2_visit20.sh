#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v20
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "entropy or simulation or quant or policy or gaussian_model" > ${P}_pytest.log 2>&1; echo "exit $?" >> ${P}_pytest.log; tail -3 ${P}_pytest.log | cut -c1-200; grep -E "^FAILED|^ERROR" ${P}_pytest.log | head
timeout 400 python bench.py --entropy --steps 20 --no-cpu-baseline --no-extra-legs > ${P}_bench.json 2> ${P}_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v20_bench.json') if l.startswith('{')][-1])
print('entropy', d['value'], d['ms_per_step'], {k:v for k,v in d['stage_ms'].items() if 'entropy' in k})
PY
