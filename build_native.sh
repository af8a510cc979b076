#!/bin/bash
# Build the in-tree native libraries (librthost.so: g++; librtgpu.so: nvcc for sm_100a).  Called by __graft_entry__.build().
#   ./build_native.sh [host|device]     RT_NVCC_EXTRA="-Xptxas -v" adds flags; RT_LIB_VARIANT=_x builds librtgpu_x.so (tuning sweeps)
set -e
cd "$(dirname "$0")"
OBJ=build/obj${RT_LIB_VARIANT}
mkdir -p rustracer_b200/lib $OBJ
NVFLAGS="-std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Iinclude $RT_NVCC_EXTRA"
D=rustracer_b200/csrc/device
if [ "$1" != "device" ]; then
  g++ -std=c++17 -O2 -ffp-contract=off -fPIC -pthread -shared -Iinclude -o rustracer_b200/lib/librthost.so rustracer_b200/csrc/host/*.cpp -lz &
fi
if [ "$1" != "host" ]; then
  pids=()
  for tu in api render tu_trace tu_rec tu_tex tu_probe bvh_build; do
    nvcc $NVFLAGS -c $D/$tu.cu -o $OBJ/$tu.o > $OBJ/$tu.log 2>&1 & pids+=($!)
  done
  for m in 0 1 2 3 4 5 6; do
    nvcc $NVFLAGS -DRT_PATH_MAT=$m -c $D/tu_path.cu -o $OBJ/tu_path_$m.o > $OBJ/tu_path_$m.log 2>&1 & pids+=($!)
  done
  fail=0
  for p in "${pids[@]}"; do wait $p || fail=1; done
  cat $OBJ/*.log
  [ $fail -eq 0 ] || { echo "nvcc failed"; exit 1; }
  nvcc -shared -o rustracer_b200/lib/librtgpu${RT_LIB_VARIANT}.so $OBJ/*.o -lcudart -ldl
fi
wait
