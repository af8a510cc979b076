#!/bin/bash
# Round-2 ncu captures (one B200).  Outputs CSVs in gpurun_out/; numbers printed under ncu are never bench values.
#   gpurun --timeout 2400 -- 'bash tools/ncu_round2.sh r02f'
TAG=${1:-r02f}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --legs none --no-cpu-baseline"
M="lts__t_bytes.sum,lts__t_sectors.sum,dram__bytes_read.sum,dram__bytes_write.sum"
# 1. every launch of the run with its device time (the timed step is the last k_raygen .. k_film_add group)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_c5.csv $B > gpurun_out/${TAG}_ncu1.out 2>&1
# 2. full capture of the closest-hit launches of one C5 step (96 per step: skip the warm-up step's)
ncu --set full --metrics $M --clock-control none -k regex:"k_trace_closest_engine|k_trace_mis_engine" -s 96 -c 96 -f -o gpurun_out/${TAG}_closest_c5 $B > gpurun_out/${TAG}_ncu2.out 2>&1
ncu -i gpurun_out/${TAG}_closest_c5.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_closest_c5_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_closest_c5.ncu-rep
# 3. the same for the C3 leg's workload
B3="python bench.py --workload c3_path --steps 1 --warmup 1 --legs none --no-cpu-baseline"
ncu --set full --metrics $M --clock-control none -k regex:"k_trace_closest_engine|k_trace_mis_engine" -s 12 -c 12 -f -o gpurun_out/${TAG}_closest_c3 $B3 > gpurun_out/${TAG}_ncu3.out 2>&1
ncu -i gpurun_out/${TAG}_closest_c3.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_closest_c3_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_closest_c3.ncu-rep
# 4. C4: the HBM-resident config (16 Mi rays per launch against 10 M triangles), second launch of each kind
ncu --set full --metrics $M --clock-control none -k regex:"k_closest_batch_engine|k_anyhit_batch_engine" -f -o gpurun_out/${TAG}_c4 python tools/c4_probe.py --reps 2 > gpurun_out/${TAG}_ncu4.out 2>&1
ncu -i gpurun_out/${TAG}_c4.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_c4_raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_c4.ncu-rep
# 5. shade kernels of the first two bounces of one C3 wave, with source
ncu --set full --clock-control none --import-source on -k regex:"k_shade_path" -s 6 -c 2 -f -o gpurun_out/${TAG}_shade_c3 $B3 > gpurun_out/${TAG}_ncu5.out 2>&1
ncu -i gpurun_out/${TAG}_shade_c3.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_full_shade_c3_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_shade_c3.ncu-rep --page source --csv > gpurun_out/${TAG}_shade_c3_source.csv 2>/dev/null; rm -f gpurun_out/${TAG}_shade_c3.ncu-rep
ls -la gpurun_out | tail -12; tail -2 gpurun_out/${TAG}_ncu4.out
