#!/bin/bash
# backward last-non-zero scan of mode A: parity, then the per-lifting-size tables (standard batch and >= 8 waves)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ALLZ=384,352,320,288,256,240,224,208,192,176,160,144,128,120,112,104,96,88,80,72,64,60,56,52,48,44,40,36,32,30,28,26,24,22,20,18,16,15,14,13,12,11,10,9,8,7,6,5,4,3,2
ROWS=17 ZCS=$ALLZ OUT=aa_zc_rows17.json timeout 900 python scripts/exp_zc.py > gpurun_out/aa_zc17.log 2>&1
WAVES=8 ROWS=17 ZCS=$ALLZ OUT=aa_zc_rows17_w8.json timeout 900 python scripts/exp_zc.py > gpurun_out/aa_zc17w8.log 2>&1
ZCS=384,320,256,240,192,128,64,16 OUT=aa_zc_allrows.json timeout 600 python scripts/exp_zc.py > gpurun_out/aa_zcall.log 2>&1
WAVES=8 ZCS=384,320,256,240,192,128,64,16 OUT=aa_zc_allrows_w8.json timeout 600 python scripts/exp_zc.py > gpurun_out/aa_zcallw8.log 2>&1
python - <<'PY'
import json
for f in ("aa_zc_rows17","aa_zc_rows17_w8","aa_zc_allrows","aa_zc_allrows_w8"):
    d=json.load(open('gpurun_out/%s.json'%f))
    v={k:round(x['g_edge_updates_per_s']) for k,x in d.items()}
    print(f, v)
    print("   min", min(v.values()), "below 800:", [k for k,x in v.items() if x<800])
PY
