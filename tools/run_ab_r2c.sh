set -x
echo "default $(SWEEP_LG=22 python tools/sweep_var.py 2>&1 | tail -1)"
for f in gpurun_ab/lib_m8m3s.so; do FORGE_EC_CUDA_LIB=$PWD/$f python tools/ab_check.py 2>&1 | tail -2; echo "$f $(FORGE_EC_CUDA_LIB=$PWD/$f SWEEP_LG=22 python tools/sweep_var.py 2>&1 | tail -1)"; done
for f in gpurun_ab/lib_ladr*.so; do FORGE_EC_GLV_LADDER=regs FORGE_EC_CUDA_LIB=$PWD/$f python tools/ab_check.py 2>&1 | tail -2; echo "$f regs $(FORGE_EC_GLV_LADDER=regs FORGE_EC_CUDA_LIB=$PWD/$f SWEEP_LG=22 python tools/sweep_var.py 2>&1 | tail -1)"; done
FORGE_EC_CUDA_LIB=$PWD/gpurun_ab/lib_ladr3.so FORGE_EC_GLV_LADDER=regs SWEEP_LG=22 SWEEP_STAGES=1 python tools/sweep_var.py 2>&1 | tail -2
SWEEP_LG=22 SWEEP_STAGES=1 python tools/sweep_var.py 2>&1 | tail -2
