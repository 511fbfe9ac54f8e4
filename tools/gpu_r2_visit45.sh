#!/bin/bash
# private-histogram tile intersection with round-robin ownership: isect / parity tests, stage timing A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v45
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > ${P}_pytest.log 2>&1; echo "exit $?" >> ${P}_pytest.log; tail -3 ${P}_pytest.log
G='radii equal|^isect|ours total'
for m in 0 1; do
  GSB200_ISECT_PRIVATE=$m timeout 150 python tools/gpu_stage_timing.py 3 > ${P}_timing_priv$m.log 2>&1
  echo "== private $m: $(grep -E "$G" ${P}_timing_priv$m.log | tr '\n' ' ')"
done
GSB200_ISECT_PRIVATE=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:isect\|replica\|tile_sort -c 40 --csv --log-file ${P}_isect_launches.csv python tools/gpu_stage_timing.py 2 > ${P}_ncu.log 2>&1; echo "ncu exit $?"; grep -E "isect|replica|tile_sort" ${P}_isect_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
