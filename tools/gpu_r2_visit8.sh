#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v8
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -4 ${P}_pytest_all.log | cut -c1-200; grep -E "^FAILED|^ERROR|headline from|grad .*outside" ${P}_pytest_all.log | head
timeout 100 python -m pytest tests/test_gpu_parity.py -q -s -p no:cacheprovider -k headline 2>&1 | grep -E "headline from|grad |passed|failed" > ${P}_headline.log; cat ${P}_headline.log
timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing.log 2>&1; grep -E "^blend|^isect|^proj|^ref_|rel err|ours total|ref  total" ${P}_timing.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_blend_stream.py tests/test_gpu_widen.py tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "long_lists or masks or photometric or split or (blend_fwd_bwd and 3-False)" > ${P}_racecheck.log 2>&1; echo "exit $?" >> ${P}_racecheck.log; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" ${P}_racecheck.log | tail -6
ONCE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:blend_stream -c 2 -o ${P}_blend_stream python tools/gpu_stage_timing.py 3 > ${P}_ncu.log 2>&1; echo "ncu exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${P}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > ${P}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
