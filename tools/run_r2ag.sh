set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 100 env SWEEP_LG=16 python tools/sweep_p256.py 2>&1 | tail -1
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lg in 22 19; do SWEEP_LG=$lg timeout 100 python tools/sweep_p256.py 2>&1 | tail -1; done
timeout 100 python tools/sweep_fixed.py 2>&1 | tail -1
for f in gpurun_ab/lib_*.so; do
  echo "== $f"
  for lg in 22 19 16; do FORGE_EC_CUDA_LIB=$PWD/$f SWEEP_LG=$lg timeout 100 python tools/sweep_p256.py 2>&1 | tail -1; done
  FORGE_EC_CUDA_LIB=$PWD/$f timeout 100 python tools/sweep_fixed.py 2>&1 | tail -1
done
