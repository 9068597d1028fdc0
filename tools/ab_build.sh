#!/bin/bash
# Development aid: build variants of libforge_ec_cuda.so that differ only in the variable-base secp256k1 kernel.
#   tools/ab_build.sh <tag> <extra nvcc flags...>   ->  gpurun_ab/lib_<tag>.so
set -e
cd "$(dirname "$0")/../forge-ec_b200/csrc"
tag=$1; shift
mkdir -p ../../gpurun_ab
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -c k_wmul_var_k1.cu -o ../build/ab_$tag.o 2> ../build/ab_$tag.log
grep -E "registers|spill" ../build/ab_$tag.log | head -2
objs=$(ls ../build/*.o | grep -v "/ab_" | grep -v k_wmul_var_k1.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_ab/lib_$tag.so $objs ../build/ab_$tag.o
