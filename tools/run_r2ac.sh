set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/sweep_x.py 2>&1 | tail -1
python tools/sweep_x.py 2>&1 | tail -1
for f in gpurun_ab/lib_*.so; do
  echo "== $f"
  FORGE_EC_CUDA_LIB=$PWD/$f python tools/sweep_x.py 2>&1 | tail -1
  FORGE_EC_CUDA_LIB=$PWD/$f python tools/sweep_x.py 2>&1 | tail -1
done
