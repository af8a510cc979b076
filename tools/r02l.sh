#!/bin/bash
# full ncu capture (with source) of the C5 shade kernels: matte and plastic, bounce 0 and 1 of the first wave of the timed step
TAG=${1:-r02l}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --legs none --no-cpu-baseline"
for m in 0 1; do
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_shade_path<\(int\)$m>" -s 48 -c 2 -f -o gpurun_out/${TAG}_shade_c5_m$m $B > gpurun_out/${TAG}_ncu_shade_m$m.out 2>&1
ncu -i gpurun_out/${TAG}_shade_c5_m$m.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_shade_c5_m${m}_raw.csv 2>/dev/null
python tools/ncu_source.py gpurun_out/${TAG}_shade_c5_m$m.ncu-rep 0 80 > gpurun_out/${TAG}_shade_c5_m${m}_bounce0_source_summary.txt 2>&1
python tools/ncu_source.py gpurun_out/${TAG}_shade_c5_m$m.ncu-rep 1 50 > gpurun_out/${TAG}_shade_c5_m${m}_bounce1_source_summary.txt 2>&1
ls -la gpurun_out/${TAG}_shade_c5_m$m.ncu-rep
[ $m = 0 ] || rm -f gpurun_out/${TAG}_shade_c5_m$m.ncu-rep
done
tail -3 gpurun_out/${TAG}_ncu_shade_m0.out
