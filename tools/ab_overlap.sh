#!/bin/bash
# A/B of the two-stream bounce overlap (rtgpu option overlap_bounces) on one box: traversal + render parity tests, then bench.py both ways, twice.
TAG=${1:-ab}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${AB_PYTEST_ARGS} > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
for rep in 1 2; do
  for ov in ${AB_VALUES:-0 1 2}; do
    RT_OPTIONS=overlap_bounces=$ov python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_ov${ov}_${rep}.json 2>> gpurun_out/${TAG}_bench.err
    python - <<PY
import json; d=json.load(open("gpurun_out/${TAG}_bench_ov${ov}_${rep}.json")); print("overlap=$ov rep=$rep", round(d["value"]/1e6,1), "M samples/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"]/1e6,1), d["rays"])
PY
  done
done
