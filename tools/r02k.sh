#!/bin/bash
# GPU tests (baseline-size parity included), shade launch-bound variants A/B, full ncu capture (with source) of the C5 shade kernels.
TAG=${1:-r02k}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for rep in 1 2; do
for v in "" _s320 _s384 _s256; do
  [ -f rustracer_b200/lib/librtgpu$v.so ] || continue
  RT_LIB_VARIANT=$v python bench.py --steps 4 --warmup 2 --legs c3_path --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); l=d['legs']['c3_path']
print('variant \"$v\" rep $rep c5', round(d['value']/1e6,1), 'shade ms', round(d['roofline_shade']['ms_per_step'],2), 'c3', round(l['value']/1e6,1), 'shade ms', round(l['roofline_shade']['ms_per_step'],2))" | tee -a gpurun_out/${TAG}_shade_variants.log
done; done
B="python bench.py --steps 1 --warmup 1 --legs none --no-cpu-baseline"
for m in 0 1; do
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_shade_path<$m>" -s 48 -c 2 -f -o gpurun_out/${TAG}_shade_c5_m$m $B > gpurun_out/${TAG}_ncu_shade_m$m.out 2>&1
ncu -i gpurun_out/${TAG}_shade_c5_m$m.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_shade_c5_m${m}_raw.csv 2>/dev/null
python tools/ncu_source.py gpurun_out/${TAG}_shade_c5_m$m.ncu-rep 0 70 > gpurun_out/${TAG}_shade_c5_m${m}_bounce0_source_summary.txt 2>&1
python tools/ncu_source.py gpurun_out/${TAG}_shade_c5_m$m.ncu-rep 1 50 > gpurun_out/${TAG}_shade_c5_m${m}_bounce1_source_summary.txt 2>&1
rm -f gpurun_out/${TAG}_shade_c5_m$m.ncu-rep
done
ls -la gpurun_out | tail -8
