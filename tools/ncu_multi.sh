#!/bin/bash
# Full ncu captures of several kernels over a short bench run.  usage: tools/ncu_multi.sh TAG "regex:skip:count ..."
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read KR SKIP CNT <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$KR -s $SKIP -c $CNT -f -o gpurun_out/${TAG}_${KR} \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_${KR}.out 2>&1
done
ls -la gpurun_out | tail -8
