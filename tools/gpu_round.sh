#!/bin/bash
# One GPU-box round: GPU tests, bench, ncu launch list and one full capture of the top kernel.  Outputs in gpurun_out/.
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 8 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_closest -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_closest \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.out 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_bench.json | cut -c1-600; ls -la gpurun_out | tail -12
