#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 python scripts/dbg_mb.py 2>&1 | tail -8
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
tail -6 gpurun_out/l_pytest.log
ROWS=17 ZCS=256,240,224,208,64 OUT=l_zc17.json timeout 300 python scripts/exp_zc.py 2>&1 | cut -c1-80
ROWS=17 ZCS=256,240,224,208 NRLDPC_NO_W8=1 OUT=l_zc17_now8.json timeout 300 python scripts/exp_zc.py 2>&1 | cut -c1-80
ROWS=12 ZCS=256,240,208 OUT=l_zc12.json timeout 300 python scripts/exp_zc.py 2>&1 | cut -c1-80
