#!/bin/bash
# bench.py C5 + C3 leg under library variants with the per-class shares: tools/ab_shares.sh TAG "" _x
TAG=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  RT_LIB_VARIANT=$v python bench.py --steps 4 --warmup 2 --legs c3_path --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); l=d['legs']['c3_path']; r=d['roofline']
print('variant \"$v\" rep $rep c5', round(d['value']/1e6,1), {k: round(x,3) for k,x in r['share_of_step'].items()}, 'closest launch ms', round(r['avg_launch_ms'],3), 'shade ms', round(d['roofline_shade']['ms_per_step'],1), '| c3', round(l['value']/1e6,1))" | tee -a gpurun_out/${TAG}_shares.log
done; done
