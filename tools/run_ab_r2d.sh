set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "k1 default $(SWEEP_LG=22 python tools/sweep_var.py 2>&1 | tail -1)"
bash tools/ab_run_p256.sh
for f in gpurun_ab/lib_p3prep256.so gpurun_ab/lib_p3prep512.so; do FORGE_EC_CUDA_LIB=$PWD/$f python tools/ab_check.py 2>&1 | tail -2; done
