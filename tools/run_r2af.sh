set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lg in 22 20 19 18 17 16; do SWEEP_LG=$lg python tools/sweep_p256.py 2>&1 | tail -1; done
for f in gpurun_ab/lib_*.so; do
  echo "== $f"
  for lg in 22 20 19 18 17 16; do FORGE_EC_CUDA_LIB=$PWD/$f SWEEP_LG=$lg python tools/sweep_p256.py 2>&1 | tail -1; done
done
