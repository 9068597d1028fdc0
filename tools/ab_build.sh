#!/bin/bash
# Development aid: build a variant of libforge_ec_cuda.so that differs in ONE translation unit.
#   tools/ab_build.sh <tag> <file.cu> <extra nvcc flags...>   ->  gpurun_ab/lib_<tag>.so
set -e
cd "$(dirname "$0")/../forge-ec_b200/csrc"
tag=$1; src=$2; shift; shift
base=${src%.cu}
mkdir -p ../../gpurun_ab
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v "$@" -c $src -o ../build/ab_$tag.o 2> ../build/ab_$tag.log
grep -E "registers|spill" ../build/ab_$tag.log | grep -B1 registers | head -8
objs=$(ls ../build/*.o | grep -v "/ab_" | grep -v "/$base.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_ab/lib_$tag.so $objs ../build/ab_$tag.o
