#!/bin/bash
# final state of the round: full suite, smoke, bench line (all legs), launch list of a bench step
mkdir -p gpurun_out
P=gpurun_out/r2v46
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -4 ${P}_pytest_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 ${P}_smoke.log
timeout 400 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v46_bench.json') if l.startswith('{')][-1])
print('N=1', d['steps'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['launch_ms'], d['clocks'])
for k in ('api_literal','quant_separate','entropy_on','ref_cuda','cpu_baseline'):
    print(k, {a:b for a,b in d.get(k,{}).items() if a in ('value','ms_per_step','unavailable','kernels_ms','cores')})
print(d['stage_ms'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ${P}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > ${P}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
