#!/bin/bash
# Final-code ncu captures for profiles/roofline_traffic.json (one B200): closest-hit launches of one C5 step and one C3 step, the C4 batch kernels.
TAG=${1:-r03k}
mkdir -p gpurun_out
M="lts__t_bytes.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum"
B="python bench.py --steps 1 --warmup 1 --legs none --no-cpu-baseline"
ncu --set full --metrics $M --clock-control none -k regex:"k_trace_closest_engine|k_trace_mis_engine" -s 96 -c 96 -f -o gpurun_out/${TAG}_closest_c5 $B > gpurun_out/${TAG}_ncu2.out 2>&1
ncu -i gpurun_out/${TAG}_closest_c5.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_closest_c5_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_closest_c5.ncu-rep
B3="python bench.py --workload c3_path --steps 1 --warmup 1 --legs none --no-cpu-baseline"
ncu --set full --metrics $M --clock-control none -k regex:"k_trace_closest_engine|k_trace_mis_engine" -s 12 -c 12 -f -o gpurun_out/${TAG}_closest_c3 $B3 > gpurun_out/${TAG}_ncu3.out 2>&1
ncu -i gpurun_out/${TAG}_closest_c3.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_closest_c3_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_closest_c3.ncu-rep
ncu --set full --metrics $M --clock-control none -k regex:"k_closest_batch_engine|k_anyhit_batch_engine" -f -o gpurun_out/${TAG}_c4 python tools/c4_probe.py --reps 2 > gpurun_out/${TAG}_ncu4.out 2>&1
ncu -i gpurun_out/${TAG}_c4.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_c4_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_c4.ncu-rep
ls -la gpurun_out | grep ${TAG}
