#!/bin/bash
# Build the in-tree native libraries (librthost.so: g++; librtgpu.so: nvcc for sm_100a).  Called by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
mkdir -p rustracer_b200/lib
NVFLAGS="-std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Iinclude"
if [ "$1" != "device" ]; then
  g++ -std=c++17 -O2 -ffp-contract=off -fPIC -pthread -shared -Iinclude -o rustracer_b200/lib/librthost.so rustracer_b200/csrc/host/*.cpp
fi
if [ "$1" != "host" ]; then
  nvcc $NVFLAGS $RT_NVCC_EXTRA -shared -o rustracer_b200/lib/librtgpu.so rustracer_b200/csrc/device/api.cu rustracer_b200/csrc/device/render.cu -lcudart
fi
