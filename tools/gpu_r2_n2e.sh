#!/bin/bash
# 2-GPU visit at the end of the round (gpurun --gpus 2): distributed parity, bench at N=2 and N=1 on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2n2e
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tools/dist_check.py > ${P}_dist_check.log 2>&1; tail -8 ${P}_dist_check.log | grep -v "^\*\|OMP"
timeout 300 $TR --master-port 29513 bench.py --gpus 2 > ${P}_bench.json 2> ${P}_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n2e_bench.json') if l.startswith('{')][-1])
print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
PY
timeout 200 python bench.py --no-cpu-baseline --no-extra-legs > ${P}_bench_n1.json 2> ${P}_bench_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n2e_bench_n1.json') if l.startswith('{')][-1])
print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
