timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "early_stop or chain" 2>&1 | tail -2
timeout 600 python scripts/exp_cfg4.py 2>&1 | tail -6
