set -x
timeout 200 python tools/sweep_msm.py 2>&1 | grep "2^22"
FORGE_EC_CUDA_LIB=$PWD/gpurun_ab/lib_nohints.so timeout 200 python tools/sweep_msm.py 2>&1 | grep "2^22"
