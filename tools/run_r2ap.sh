set -x
timeout 100 python tools/sweep_fixed.py 2>&1 | tail -1
for f in gpurun_ab/lib_*.so; do
  echo "== $f"
  FORGE_EC_CUDA_LIB=$PWD/$f timeout 100 python tools/sweep_fixed.py 2>&1 | tail -1
done
