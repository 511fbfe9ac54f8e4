#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_n2.log 2>&1; echo "dist_check exit $?"; tail -12 gpurun_out/dist_check_n2.log
timeout 200 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
