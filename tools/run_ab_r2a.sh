set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for f in gpurun_ab/lib_dual4.so gpurun_ab/lib_p3dual4.so; do FORGE_EC_CUDA_LIB=$PWD/$f python tools/ab_check.py; done
echo "default $(SWEEP_LG=22 python tools/sweep_var.py 2>&1 | tail -1)"
bash tools/ab_run.sh
bash tools/ab_run_p256.sh
