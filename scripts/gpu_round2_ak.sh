#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kernel_variants or every_lifting or multi_block" 2>&1 | tail -3
