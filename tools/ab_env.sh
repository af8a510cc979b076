#!/bin/bash
# bench.py (C5 + C3 + AO + C4 legs) under settings of one environment variable: tools/ab_env.sh TAG VAR v1 v2 ...
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  env $VAR=$v python bench.py --steps 4 --warmup 2 --legs c3_path,c3_ao,c4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); L=d['legs']; l=L['c3_path']
print('$VAR=$v rep $rep c5', round(d['value']/1e6,1), '| c3', round(l['value']/1e6,1), '| ao Mrays/s', round(L['c3_ao']['mrays_per_s'],1), '| c4 closest', round(L['c4']['closest']['mrays_per_s_device'],1), 'anyhit', round(L['c4']['anyhit']['mrays_per_s_device'],1))" | tee -a gpurun_out/${TAG}_env.log
done; done
