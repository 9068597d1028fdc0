set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
SWEEP_LG=22 SWEEP_STAGES=1 python tools/sweep_var.py 2>&1 | tail -2
for f in gpurun_ab/lib_*.so; do
  echo "== $f"
  FORGE_EC_CUDA_LIB=$PWD/$f SWEEP_LG=22 SWEEP_STAGES=1 python tools/sweep_var.py 2>&1 | tail -2
done
python tools/sweep_p256.py 2>&1 | tail -1
