#!/bin/bash
# one-wave jobs (what a rank of an 8-GPU run renders per step): one 32 M-path wave against two 16 M-path waves in flight
TAG=${1:-r03e}
mkdir -p gpurun_out
for rep in 1 2; do
for wp in 0 16777216 11184811 8388608; do
  RT_WAVE_PATHS=$wp python bench.py --steps 8 --warmup 3 --spp-per-step 4 --legs none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('wave_paths=$wp rep $rep c5 4 spp per step:', round(d['value']/1e6,1), 'M samples/s', round(d['ms_per_step'],2), 'ms/step, launches', d['gpu_launches'])" | tee -a gpurun_out/${TAG}_wave_split.log
done; done
