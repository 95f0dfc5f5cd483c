#!/bin/bash
# Experiment helper: builds ../variants/libkprn_<name>.so with extra -D flags for tc_path.cu (other objects are shared).
# usage: build_variants.sh name1:"-DA=1 -DB=1" name2:"..."       (select at run time with KPRN_B200_LIB=<path>)
set -e
cd "$(dirname "$0")"
make -s
mkdir -p ../variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ -ftz=true"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( $NV $flags -Xptxas -v -c tc_path.cu -o ../variants/tc_path_$name.o 2> ../variants/ptxas_$name.log
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libkprn_$name.so api.o kernels_simple.o tc_gemm.o ../variants/tc_path_$name.o -cudart static
    echo "built $name ($flags)" ) &
done
wait
