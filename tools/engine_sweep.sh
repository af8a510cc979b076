#!/bin/bash
# GPU-box sweep of the traversal engine: parity first, then scheduling knobs for each built library variant.
#   gpurun -- 'bash tools/engine_sweep.sh TAG "_mb6 _x" [tune args]'
TAG=${1:-sweep}; VARIANTS=${2:-}; shift 2
mkdir -p gpurun_out
python -m pytest tests/test_gpu_traversal.py tests/test_gpu_render.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python tools/tune_engine.py "$@" > gpurun_out/${TAG}_tune.log 2>&1
for v in $VARIANTS; do RT_LIB_VARIANT=$v python tools/tune_engine.py "$@" > gpurun_out/${TAG}_tune$v.log 2>&1; done
cat gpurun_out/${TAG}_tune*.log
