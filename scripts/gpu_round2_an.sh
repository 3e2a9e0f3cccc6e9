#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
ROWS=17 ZCS=128,120,112,104,26 OUT=an_a$rep.json timeout 300 python scripts/exp_zc.py 2>&1 | grep BG | awk '{print $1,$2,$8}' | tr '\n' ';'; echo
WAVES=8 ROWS=17 ZCS=384,9 OUT=an_b$rep.json timeout 300 python scripts/exp_zc.py 2>&1 | grep BG | awk '{print $1,$2,$8}' | tr '\n' ';'; echo
done
