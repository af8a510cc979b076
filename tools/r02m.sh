#!/bin/bash
# sampler closed forms parity check, shade block-size variants, small-batch latency table, source-level capture of the closest-hit engine (C5 bounce 1)
TAG=${1:-r02m}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_render.py tests/test_gpu_pins.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for rep in 1 2; do
for v in "" _t128 _t256; do
  [ -f rustracer_b200/lib/librtgpu$v.so ] || continue
  RT_LIB_VARIANT=$v python bench.py --steps 4 --warmup 2 --legs c3_path --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); l=d['legs']['c3_path']
print('variant \"$v\" rep $rep c5', round(d['value']/1e6,1), 'shade ms', round(d['roofline_shade']['ms_per_step'],2), 'c3', round(l['value']/1e6,1), 'shade ms', round(l['roofline_shade']['ms_per_step'],2))" | tee -a gpurun_out/${TAG}_shade_variants.log
done; done
python tools/small_batches.py > gpurun_out/${TAG}_small_batches.log 2>&1; cat gpurun_out/${TAG}_small_batches.log
B="python bench.py --steps 1 --warmup 1 --legs none --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest_engine" -s 97 -c 1 -f -o gpurun_out/${TAG}_closest_c5_b1 $B > gpurun_out/${TAG}_ncu_closest.out 2>&1
python tools/ncu_source.py gpurun_out/${TAG}_closest_c5_b1.ncu-rep 0 70 > gpurun_out/${TAG}_closest_c5_bounce1_source_summary.txt 2>&1
ncu -i gpurun_out/${TAG}_closest_c5_b1.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_closest_c5_bounce1_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_closest_c5_b1.ncu-rep
head -50 gpurun_out/${TAG}_closest_c5_bounce1_source_summary.txt
