timeout 600 python -m pytest tests/test_gpu_sweep.py -x -q 2>&1 | tail -2
timeout 600 python scripts/exp_cfg4.py 2>&1 | tail -6
