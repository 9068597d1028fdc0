#!/bin/bash
# Development aid: time every gpurun_ab/lib_*.so variant of the k*P kernel (run on the GPU box).
for f in gpurun_ab/lib_*.so; do
  echo -n "$f  "
  FORGE_EC_CUDA_LIB=$PWD/$f SWEEP_LG=${SWEEP_LG:-22} python tools/sweep_var.py 2>&1 | tail -1
done
