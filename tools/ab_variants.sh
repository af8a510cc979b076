#!/bin/bash
# bench.py (C5 + C3 leg) under library variants on one box: tools/ab_variants.sh TAG "" _x _y
TAG=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  RT_LIB_VARIANT=$v python bench.py --steps 4 --warmup 2 --legs c3_path --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); l=d['legs']['c3_path']
print('variant \"$v\" rep $rep c5', round(d['value']/1e6,1), 'shade ms', round(d['roofline_shade']['ms_per_step'],2), 'c3', round(l['value']/1e6,1), 'shade ms', round(l['roofline_shade']['ms_per_step'],2))" | tee -a gpurun_out/${TAG}_variants.log
done; done
