#!/bin/bash
# after the shared-memory budget fix and the tiered kernels without early-termination code: full suite, tables, quick A/B figures
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ALLZ=384,352,320,288,256,240,224,208,192,176,160,144,128,120,112,104,96,88,80,72,64,60,56,52,48,44,40,36,32,30,28,26,24,22,20,18,16,15,14,13,12,11,10,9,8,7,6,5,4,3,2
ROWS=17 ZCS=$ALLZ OUT=aa_zc_rows17.json timeout 900 python scripts/exp_zc.py > gpurun_out/aa_zc17.log 2>&1
WAVES=8 ROWS=17 ZCS=$ALLZ OUT=aa_zc_rows17_w8.json timeout 900 python scripts/exp_zc.py > gpurun_out/aa_zc17w8.log 2>&1
AZ=384,352,320,288,256,240,224,208,192,128,64,16
ZCS=$AZ OUT=aa_zc_allrows.json timeout 600 python scripts/exp_zc.py > gpurun_out/aa_zcall.log 2>&1
WAVES=8 ZCS=$AZ OUT=aa_zc_allrows_w8.json timeout 600 python scripts/exp_zc.py > gpurun_out/aa_zcallw8.log 2>&1
WAVES=8 ROWS=30 ZCS=$AZ OUT=aa_zc_rows30_w8.json timeout 600 python scripts/exp_zc.py > /dev/null 2>&1
python - <<'PY'
import json
for f in ("aa_zc_rows17","aa_zc_rows17_w8","aa_zc_allrows","aa_zc_allrows_w8","aa_zc_rows30_w8"):
    d=json.load(open('gpurun_out/%s.json'%f))
    v={k:round(x['g_edge_updates_per_s']) for k,x in d.items()}
    print(f, v if len(v)<30 else '', "min", min(v.values()), "below 800:", [k for k,x in v.items() if x<800])
PY
rm -f gpurun_out/am_ab.jsonl
for r in 0.6 0.45 0.4 0.33; do timeout 200 python scripts/ab_quick.py --tbs 64 --steps 40 --rate $r --tag "r$r" >> gpurun_out/am_ab.jsonl 2>/dev/null; done
python - <<'PY'
import json
for l in open('gpurun_out/am_ab.jsonl'):
    d=json.loads(l); print(d['tag'], d['single_gbps'], d['two_stream_gbps'], d['tb_ok'], d['bit_err'])
PY
