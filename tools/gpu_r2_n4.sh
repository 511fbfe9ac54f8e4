#!/bin/bash
# N-GPU visit (gpurun --gpus N, N passed as $1): distributed parity check and bench at N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
P=gpurun_out/r2n${N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tools/dist_check.py > ${P}_dist_check.log 2>&1; tail -4 ${P}_dist_check.log | grep -v "^\*\|OMP"
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; python - <<PY
import json
d=json.loads([l for l in open('${P}_bench.json') if l.startswith('{')][-1])
print('N=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
PY
tail -3 ${P}_bench.err
