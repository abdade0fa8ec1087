#!/bin/bash
# developer tool: builds named variants of the library into build/ (tools/gpu_*.sh select them with SWB_LIB)
#   tools/build_variants.sh name1 "-DFLAG ..." name2 "-DFLAG ..." ...
cd "$(dirname "$0")/.."
mkdir -p build
SRC="smith-waterman_b200/csrc/swb_api.cu smith-waterman_b200/csrc/swb_multi.cu smith-waterman_b200/csrc/swb_pack.cu smith-waterman_b200/csrc/swb_io.cpp"
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ccbin /usr/bin/g++ -shared"
while [ $# -ge 2 ]; do
  ( $NV $2 -o build/libswb200_$1.so $SRC -lcudart 2>&1 | grep -E "error" ; echo "built $1 ($2)" ) &
  shift 2
done
wait
