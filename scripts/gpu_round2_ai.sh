#!/bin/bash
# three-CTA tiered kernels for CTAs of at most 8 warps at low code rates: parity, then all rows / 36 / 30 rows at Zc = 208 .. 256
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_lifting or kernel_variants or multi_block or small_z" 2>&1 | tail -3
Z=256,240,224,208
WAVES=8 ZCS=$Z OUT=ai_all_new.json timeout 300 python scripts/exp_zc.py > /dev/null 2>&1
NRLDPC_W8_TIERED_FROM=99 WAVES=8 ZCS=$Z OUT=ai_all_old.json timeout 300 python scripts/exp_zc.py > /dev/null 2>&1
for r in 36 30; do
NRLDPC_W8_TIERED_FROM=27 WAVES=8 ROWS=$r ZCS=$Z OUT=ai_r${r}_new.json timeout 300 python scripts/exp_zc.py > /dev/null 2>&1
NRLDPC_W8_TIERED_FROM=99 WAVES=8 ROWS=$r ZCS=$Z OUT=ai_r${r}_old.json timeout 300 python scripts/exp_zc.py > /dev/null 2>&1
done
python - <<'PY'
import json
for tag in ("all","r36","r30"):
    n=json.load(open('gpurun_out/ai_%s_new.json'%tag)); o=json.load(open('gpurun_out/ai_%s_old.json'%tag))
    print(tag, {k:(round(n[k]['g_edge_updates_per_s']), round(o[k]['g_edge_updates_per_s'])) for k in n})
PY
