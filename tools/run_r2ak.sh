set -x
timeout 200 python tools/sweep_msm.py 2>&1 | grep "2^22\|2^18"
timeout 300 python -m pytest tests -m gpu -x -q -k "msm" 2>&1 | tail -3
