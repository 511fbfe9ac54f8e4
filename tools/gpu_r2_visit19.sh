#!/bin/bash
# entropy-on step: stage times and launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v19
timeout 400 python bench.py --entropy --steps 10 --no-cpu-baseline --no-extra-legs > ${P}_bench.json 2> ${P}_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2v19_bench.json') if l.startswith('{')][-1])
print('entropy', d['value'], d['ms_per_step'], d['stage_ms'])
PY
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread --clock-control none -c 500 --csv --log-file ${P}_launches.csv python bench.py --entropy --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > ${P}_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2v19_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); im=hdr.index('Metric Name'); iid=hdr.index('ID')
L={}
order=[]
for r in rows[1:]:
    if r[iid] not in L: L[r[iid]]={'k':r[ik]}; order.append(r[iid])
    L[r[iid]][r[im]]=r[iv]
seq=[L[i] for i in order]
st=[i for i,e in enumerate(seq) if 'fake_quantize_multi_fwd' in e['k']]
for e in seq[st[-2]:st[-1]]:
    t=float(e.get('gpu__time_duration.sum','0').replace(',',''))/1000
    if t>8: print(f"{t:8.1f} us inst {e.get('smsp__inst_executed.sum')} sm% {e.get('sm__throughput.avg.pct_of_peak_sustained_elapsed')} regs {e.get('launch__registers_per_thread')}  {e['k'][:90]}")
PY
