timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/exp_cfg4.py 2>&1 | tail -6
