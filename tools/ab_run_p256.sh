#!/bin/bash
# Development aid: time every gpurun_ab/lib_*.so variant of the P-256 k*P path (run on the GPU box).
echo -n "default  "; python tools/sweep_p256.py 2>&1 | tail -1
echo -n "mono     "; FORGE_EC_VAR_IMPL=mono python tools/sweep_p256.py 2>&1 | tail -1
for f in gpurun_ab/lib_*.so; do
  echo -n "$f  "
  FORGE_EC_CUDA_LIB=$PWD/$f python tools/sweep_p256.py 2>&1 | tail -1
done
