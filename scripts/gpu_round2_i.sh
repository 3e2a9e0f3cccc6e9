#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python scripts/dbg_mb.py 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=240 > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
tail -15 gpurun_out/i_pytest.log
ROWS=17 ZCS=240,208,192,176,144,128,96,64,36,32,16,8,2 OUT=r02_zc_rows17_mb.json timeout 600 python scripts/exp_zc.py 2>&1 | cut -c1-80
