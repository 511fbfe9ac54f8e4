#!/bin/bash
# depth-bucket sort for long segments too: isect tests, timing, launch list, ncu --set full of the two sort kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v12
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "isect or golden or full_size or headline" > ${P}_isect_tests.log 2>&1; echo "exit $?" >> ${P}_isect_tests.log; tail -3 ${P}_isect_tests.log | cut -c1-200
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; grep -E "^isect|^ref_isect|ours total" ${P}_timing.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"isect_|tile_sort|replica_|scan_" -c 20 --csv --log-file ${P}_isect_launches.csv python tools/gpu_stage_timing.py 3 > /dev/null 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2v12_isect_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[1:11]: print(r[ik][:50], r[iv])
PY
ONCE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tile_sort_bucket -c 2 -o ${P}_tile_sort python tools/gpu_stage_timing.py 3 > ${P}_ncu.log 2>&1; ls -la ${P}_tile_sort.ncu-rep
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -3 ${P}_pytest_all.log | cut -c1-200; grep -E "^FAILED|^ERROR" ${P}_pytest_all.log | head
