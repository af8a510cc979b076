#!/bin/bash
# Round-end evidence on one box: GPU tests, smoke, default bench (all legs + CPU baseline), reference arm, C1 / C2 configs beside the CPU oracle,
# launch list of one C5 step
TAG=${1:-r02p}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -4 gpurun_out/${TAG}_bench.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -4 gpurun_out/${TAG}_bench_reference.err
python tools/run_configs.py --configs c1,c2 --out gpurun_out/${TAG}_configs_c1_c2.json > gpurun_out/${TAG}_configs.log 2>&1; tail -3 gpurun_out/${TAG}_configs.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_c5.csv python bench.py --steps 1 --warmup 1 --legs none --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.out 2>&1
python - $TAG <<'P'
import json,sys
d=json.loads(open('gpurun_out/'+sys.argv[1]+'_bench.json' if len(sys.argv)>1 else 'gpurun_out/r02p_bench.json').read().strip().splitlines()[-1])
print(d['value']/1e6, d['e2e']['value']/1e6, d.get('cpu_baseline'), {k:(v.get('value') or v.get('closest',{}).get('mrays_per_s_device')) for k,v in d['legs'].items()})
P
