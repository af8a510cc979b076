#!/bin/bash
# bench.py under several RT_OPTIONS settings on one box: tools/ab_options.sh TAG "opt=v,opt=v" "..." ...
TAG=$1; shift
mkdir -p gpurun_out
: > gpurun_out/${TAG}_ab_options.log
for rep in 1 2; do
  for o in "$@"; do
    RT_OPTIONS=$o python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$o rep $rep', round(d['value']/1e6,1), 'M samples/s', round(d['ms_per_step'],3), 'ms/step')" | tee -a gpurun_out/${TAG}_ab_options.log
  done
done
