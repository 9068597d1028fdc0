set -x
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/sweep_msm.py 2>&1 | tail -4
