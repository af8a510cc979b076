#!/bin/bash
# diagonal tile dealing: tile tests + per-rank balance on one GPU; any-hit order mode 2 (near pair first) against mode 1 (storage order)
TAG=${1:-r02q}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "tile or sample_ranges or li_and_image or baseline" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python tools/tile_balance.py 2,4,8 > gpurun_out/${TAG}_tile_balance.log 2>&1; cat gpurun_out/${TAG}_tile_balance.log
for rep in 1 2; do
for v in "" _afo2; do
  RT_LIB_VARIANT=$v python bench.py --steps 4 --warmup 2 --legs c3_path,c3_ao --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); L=d['legs']; l=L['c3_path']
print('variant \"$v\" rep $rep c5', round(d['value']/1e6,1), 'anyhit share', round(d['roofline']['share_of_step']['anyhit'],3), '| c3', round(l['value']/1e6,1), 'anyhit share', round(l['roofline']['share_of_step']['anyhit'],3),
 '| ao Mrays/s', round(L['c3_ao']['mrays_per_s'],1))" | tee -a gpurun_out/${TAG}_ab_afo2.log
done; done
