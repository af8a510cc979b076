#!/bin/bash
# ncu launch list + full capture of one kernel over a short bench run.  usage: tools/ncu_round.sh TAG KERNEL_REGEX [skip] [count]
TAG=$1; KR=${2:-k_trace_closest}; SKIP=${3:-4}; CNT=${4:-2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KR -s $SKIP -c $CNT -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.out 2>&1
ls -la gpurun_out | tail -5
