#!/bin/bash
# variant parity (whole GPU suite under RT_LIB_VARIANT) + A/B against the default build
TAG=${1:-r02s}; V=$2
mkdir -p gpurun_out
RT_LIB_VARIANT=$V python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest$V.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest$V.log
tail -4 gpurun_out/${TAG}_pytest$V.log
tools/ab_variants.sh $TAG "" $V
