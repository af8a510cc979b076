#!/bin/bash
# strong scaling of the C5 job on one box: N = 8 and N = 4 through the driver's launch line (each also renders the job on rank 0 alone for the film check)
TAG=${1:-r02o}
mkdir -p gpurun_out
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
  echo "n=$n rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_n$n.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','n_gpus','ms_per_step','scaling','reduce_ms','gpu_launches')}); print(d.get('film_check')); print(d.get('e2e'))"
done
tail -5 gpurun_out/${TAG}_bench_n8.err
