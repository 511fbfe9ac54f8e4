#!/bin/bash
# One visit to the GPU box (`gpurun -- 'bash tools/gpu_visit.sh [steps...]'`): most important step first, every step
# bounded by its own timeout, everything written under gpurun_out/.  Steps: tests smoke bench bench-cat timing launches
# ncu-loss n2 (n2 needs `gpurun --gpus 2`).  Default: tests smoke bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
steps=${@:-tests smoke bench}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for s in $steps; do
  case $s in
    tests)     timeout 400 python -m pytest tests -m gpu -q -x --durations=6 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
               echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log ;;
    smoke)     timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
               echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log ;;
    bench)     timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json ;;
    bench-cat) timeout 100 python bench.py --cat-colors --torch-loss --no-cpu-baseline > gpurun_out/bench_cat.json 2> gpurun_out/bench_cat.err
               cat gpurun_out/bench_cat.json ;;
    timing)    timeout 100 python tools/widen_timing.py > gpurun_out/widen_timing.log 2>&1; cat gpurun_out/widen_timing.log
               timeout 150 python tools/gpu_stage_timing.py 3 > gpurun_out/stage_timing.log 2>&1; tail -20 gpurun_out/stage_timing.log ;;
    launches)  timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
                 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?" ;;
    ncu-loss)  timeout 150 ncu --set full --clock-control none --import-source on -k regex:"photometric" -s 6 -c 2 -o gpurun_out/prof_loss \
                 python tools/widen_timing.py > gpurun_out/ncu_loss.log 2>&1; echo "ncu exit $?" ;;
    n2)        timeout 150 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_n2.log 2>&1; tail -8 gpurun_out/dist_check_n2.log
               timeout 200 $TR --master-port 29512 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json ;;
    *)         echo "unknown step $s" ;;
  esac
done
