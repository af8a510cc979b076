#!/bin/bash
# any-hit walks in storage order (RT_ENGINE_ANY_FIXED_ORDER): parity under the variant, then A/B on C5 / C3 path, C3 AO and the C4 batches
TAG=${1:-r02n}; V=${2:-_afo}
mkdir -p gpurun_out
RT_LIB_VARIANT=$V python -m pytest tests/test_gpu_traversal.py tests/test_gpu_render.py tests/test_ray_batch.py -m gpu -x -q > gpurun_out/${TAG}_pytest$V.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest$V.log
tail -4 gpurun_out/${TAG}_pytest$V.log
for rep in 1 2; do
for v in "" $V; do
  RT_LIB_VARIANT=$v python bench.py --steps 4 --warmup 2 --legs c3_path,c3_ao,c4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); L=d['legs']; l=L['c3_path']
print('variant \"$v\" rep $rep c5', round(d['value']/1e6,1), 'anyhit share', round(d['roofline']['share_of_step']['anyhit'],3), '| c3', round(l['value']/1e6,1), 'anyhit share', round(l['roofline']['share_of_step']['anyhit'],3),
 '| ao Mrays/s', round(L['c3_ao']['mrays_per_s'],1), '| c4 closest', round(L['c4']['closest']['mrays_per_s_device'],1), 'anyhit', round(L['c4']['anyhit']['mrays_per_s_device'],1))" | tee -a gpurun_out/${TAG}_ab$V.log
done; done
