#!/bin/bash
# BASELINE configs[4]: one BLER sweep of three 10^6-block points on the final tree
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/bler_sweep.py --tbs 250000 --bg 1 --mod 16QAM --A 33672 --rate 0.6 --iters 8 --snrs=8.4,8.6,9.0 --batch-tbs 2048 2>/dev/null | tee gpurun_out/ap_sweep.jsonl | cut -c1-400 | head -5
