#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v5
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "isect or golden or full_size" > ${P}_isect_tests.log 2>&1; echo "exit $?" >> ${P}_isect_tests.log; tail -15 ${P}_isect_tests.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "isect and not full" > ${P}_isect_memcheck.log 2>&1; echo "exit $?" >> ${P}_isect_memcheck.log; tail -6 ${P}_isect_memcheck.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -12 ${P}_pytest_all.log
GSB200_ISECT=ordered timeout 300 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "isect" > ${P}_tilesort_tests.log 2>&1; echo "exit $?" >> ${P}_tilesort_tests.log; tail -3 ${P}_tilesort_tests.log
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; grep -E "^isect|^blend|^ref_isect|ours total" ${P}_timing.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vis_|radix_|piece_|cam_pieces|scan_" -c 60 --csv --log-file ${P}_isect_launches.csv python tools/gpu_stage_timing.py 3 > /dev/null 2>&1; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2v5_isect_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[1:40]: print(r[ik][:60], r[iv])
PY
