#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 python scripts/dbg_mb.py 2>&1 | tail -32
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
tail -12 gpurun_out/j_pytest.log
