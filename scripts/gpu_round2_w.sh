#!/bin/bash
# tiered multi-block static kernels (low code rates at Zc <= 192 and the lifting sizes that are no multiple of 32)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 python scripts/dbg_mb.py 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_block or every_lifting or kernel_variants" 2>&1 | tail -3
ZCS=192,176,128,96,64,36,16,208,240 OUT=w_allrows_new.json timeout 300 python scripts/exp_zc.py 2>&1 | grep BG > gpurun_out/w_new.txt
NRLDPC_NO_STATIC_MB_TIERED=1 ZCS=192,176,128,96,64,36,16,208,240 OUT=w_allrows_old.json timeout 300 python scripts/exp_zc.py 2>&1 | grep BG > gpurun_out/w_old.txt
ROWS=30 ZCS=192,128,64,16,208 OUT=w_r30_new.json timeout 300 python scripts/exp_zc.py 2>&1 | grep BG > gpurun_out/w_r30_new.txt
NRLDPC_NO_STATIC_MB_TIERED=1 ROWS=30 ZCS=192,128,64,16,208 OUT=w_r30_old.json timeout 300 python scripts/exp_zc.py 2>&1 | grep BG > gpurun_out/w_r30_old.txt
paste gpurun_out/w_new.txt gpurun_out/w_old.txt | awk '{print $1,$2,$8,"G new |",$19,"G old"}'
echo rows30; paste gpurun_out/w_r30_new.txt gpurun_out/w_r30_old.txt | awk '{print $1,$2,$8,"G new |",$19,"G old"}'
