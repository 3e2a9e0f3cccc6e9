#!/bin/bash
# helper kernels: full GPU suite, then timing (16 384 and 4 096 code blocks)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_round2_u.sh
