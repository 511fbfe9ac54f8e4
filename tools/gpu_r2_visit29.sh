#!/bin/bash
# resident-CTA cap sweep for the stream blend kernels (long-tile tail)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2v29
for n in 0 6 5 4 3 2; do
  if [ $n = 0 ]; then unset GSB200_SFWD_CTAS GSB200_SBWD_CTAS; else export GSB200_SFWD_CTAS=$n GSB200_SBWD_CTAS=$n; fi
  timeout 200 python tools/gpu_stage_timing.py 3 > ${P}_timing_$n.log 2>&1; echo "CTAs/SM cap $n: $(grep '^blend_fwd\|^blend_bwd' ${P}_timing_$n.log | tr '\n' ' ')"
done
