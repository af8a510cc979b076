#!/bin/bash
# Two-rank plain-C NCCL test with NCCL's own log, system libnccl and torch's bundled one (run under gpurun --gpus 2).
mkdir -p gpurun_out
TAG=${1:-r02j}
timeout 400 python -m pytest tests/test_c_abi.py -m gpu -q 2>&1 | tail -15 > gpurun_out/${TAG}_c_abi_pytest.log
cat gpurun_out/${TAG}_c_abi_pytest.log
NCCL_DEBUG=INFO timeout 150 build/abi_smoke 2 > gpurun_out/${TAG}_abi2_sysnccl.log 2>&1; echo "system nccl rc=$?"
tail -5 gpurun_out/${TAG}_abi2_sysnccl.log
LD_LIBRARY_PATH=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/lib:$LD_LIBRARY_PATH NCCL_DEBUG=INFO timeout 150 build/abi_smoke 2 > gpurun_out/${TAG}_abi2_torchnccl.log 2>&1; echo "torch nccl rc=$?"
tail -5 gpurun_out/${TAG}_abi2_torchnccl.log
