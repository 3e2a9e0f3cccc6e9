#!/bin/bash
# per-lifting-size table at >= 8 full waves (steady state), 17 scheduled rows and all rows
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ALLZ=384,352,320,288,256,240,224,208,192,176,160,144,128,120,112,104,96,88,80,72,64,60,56,52,48,44,40,36,32,30,28,26,24,22,20,18,16,15,14,13,12,11,10,9,8,7,6,5,4,3,2
WAVES=8 ROWS=17 ZCS=$ALLZ OUT=x_zc_rows17_w8.json timeout 900 python scripts/exp_zc.py > gpurun_out/x_zc17.log 2>&1; grep -c BG gpurun_out/x_zc17.log
WAVES=8 ZCS=384,320,256,240,192,128,64,16 OUT=x_zc_allrows_w8.json timeout 600 python scripts/exp_zc.py > gpurun_out/x_zcall.log 2>&1; grep -c BG gpurun_out/x_zcall.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/x_zc_rows17_w8.json'))
v={k:round(x['g_edge_updates_per_s']) for k,x in d.items()}
print(v)
print("min", min(v.values()), "below 800:", [k for k,x in v.items() if x<800])
print({k:round(x['g_edge_updates_per_s']) for k,x in json.load(open('gpurun_out/x_zc_allrows_w8.json')).items()})
PY
