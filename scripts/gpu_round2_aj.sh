#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
Z=256,240,224,208
for r in 0 42 38 34; do
WAVES=8 ROWS=$r ZCS=$Z OUT=aj_r${r}_new.json timeout 300 python scripts/exp_zc.py > /dev/null 2>&1
NRLDPC_NO_W8_TIERED_OCC1=1 WAVES=8 ROWS=$r ZCS=$Z OUT=aj_r${r}_old.json timeout 300 python scripts/exp_zc.py > /dev/null 2>&1
done
python - <<'PY'
import json
for tag in ("r0","r42","r38","r34"):
    n=json.load(open('gpurun_out/aj_%s_new.json'%tag)); o=json.load(open('gpurun_out/aj_%s_old.json'%tag))
    print(tag, {k:(round(n[k]['g_edge_updates_per_s']), round(o[k]['g_edge_updates_per_s'])) for k in n})
PY
