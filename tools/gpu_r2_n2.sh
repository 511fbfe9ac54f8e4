#!/bin/bash
# 2-GPU visit (gpurun --gpus 2): distributed parity check, timeline of one step, bench at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2n2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tools/dist_check.py > ${P}_dist_check.log 2>&1; tail -12 ${P}_dist_check.log
timeout 300 $TR --master-port 29512 tools/dist_profile.py > ${P}_dist_profile.log 2>&1; head -3 ${P}_dist_profile.log; grep -c . ${P}_dist_profile.log
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; cat ${P}_bench.json | cut -c1-1500; tail -3 ${P}_bench.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-legs > ${P}_bench_n1.json 2> ${P}_bench_n1.err; cut -c1-400 ${P}_bench_n1.json
