set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SWEEP_LG=22 SWEEP_STAGES=1 python tools/sweep_var.py 2>&1 | tail -2
python tools/sweep_e2e.py 2>&1 | tail -4
