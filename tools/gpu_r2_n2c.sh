#!/bin/bash
# 2-GPU visit: whole GPU suite, distributed parity check, bench at N=2 and N=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2n2d
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -3 ${P}_pytest_all.log | cut -c1-200; grep -E "^FAILED|^ERROR" ${P}_pytest_all.log | head
timeout 200 $TR --master-port 29511 tools/dist_check.py > ${P}_dist_check.log 2>&1; tail -3 ${P}_dist_check.log | grep -v "^\*\|OMP"
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n2d_bench.json') if l.startswith('{')][-1])
print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
PY
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs > ${P}_bench_n1.json 2> ${P}_bench_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n2d_bench_n1.json') if l.startswith('{')][-1])
print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['launch_ms'])
PY
