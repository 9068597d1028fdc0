set -x
python tools/sweep_fixed.py 2>&1 | tail -1
SWEEP_LG=22 SWEEP_STAGES=1 python tools/sweep_var.py 2>&1 | tail -2
python tools/sweep_p256.py 2>&1 | tail -1
for f in gpurun_ab/lib_*.so; do
  echo "== $f"
  FORGE_EC_CUDA_LIB=$PWD/$f python tools/sweep_fixed.py 2>&1 | tail -1
  FORGE_EC_CUDA_LIB=$PWD/$f SWEEP_LG=22 SWEEP_STAGES=1 python tools/sweep_var.py 2>&1 | tail -2
  FORGE_EC_CUDA_LIB=$PWD/$f python tools/sweep_p256.py 2>&1 | tail -1
  FORGE_EC_CUDA_LIB=$PWD/$f SWEEP_LG=18 python tools/sweep_fixed.py 2>&1 | tail -1
done
SWEEP_LG=18 python tools/sweep_fixed.py 2>&1 | tail -1
