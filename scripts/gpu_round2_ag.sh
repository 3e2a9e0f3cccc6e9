#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_link.py -m gpu -x -q -k fused_symbol 2>&1 | grep -v "^$" | tail -40
