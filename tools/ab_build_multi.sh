#!/bin/bash
# Development aid: a variant of libforge_ec_cuda.so that differs in SEVERAL translation units.
#   tools/ab_build_multi.sh <tag> "<a.cu b.cu ...>" <extra nvcc flags...>   ->  gpurun_ab/lib_<tag>.so
set -e
cd "$(dirname "$0")/../forge-ec_b200/csrc"
tag=$1; srcs=$2; shift; shift
mkdir -p ../../gpurun_ab
objs=$(ls ../build/*.o | grep -v "/ab_")
for src in $srcs; do
  base=${src%.cu}
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c $src -o ../build/ab_${tag}_$base.o &
  objs=$(echo "$objs" | grep -v "/$base.o")
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_ab/lib_$tag.so $objs ../build/ab_${tag}_*.o
ls -la ../../gpurun_ab/lib_$tag.so
