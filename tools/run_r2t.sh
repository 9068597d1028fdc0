set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/sweep_fixed.py 2>&1 | tail -1
FORGE_EC_CT_TABLE=1 python tools/sweep_fixed.py 2>&1 | tail -1
