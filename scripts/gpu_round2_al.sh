#!/bin/bash
# tiered one-block kernels without the early-termination code (all rows at Zc = 288 .. 384) against the ones with it
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kernel_variants or every_lifting" 2>&1 | tail -2
Z=384,352,320,288
WAVES=8 ZCS=$Z OUT=al_all_new.json timeout 300 python scripts/exp_zc.py > /dev/null 2>&1
NRLDPC_TIERED_ES_CODE=1 WAVES=8 ZCS=$Z OUT=al_all_old.json timeout 300 python scripts/exp_zc.py > /dev/null 2>&1
timeout 200 python scripts/ab_quick.py --tbs 64 --steps 40 --rate 0.33 --tag "r033 new" > gpurun_out/al_ab.jsonl 2>/dev/null
NRLDPC_TIERED_ES_CODE=1 timeout 200 python scripts/ab_quick.py --tbs 64 --steps 40 --rate 0.33 --tag "r033 old" >> gpurun_out/al_ab.jsonl 2>/dev/null
python - <<'PY'
import json
n=json.load(open('gpurun_out/al_all_new.json')); o=json.load(open('gpurun_out/al_all_old.json'))
print({k:(round(n[k]['g_edge_updates_per_s']), round(o[k]['g_edge_updates_per_s'])) for k in n})
for l in open('gpurun_out/al_ab.jsonl'):
    d=json.loads(l); print(d['tag'], d['single_gbps'], d['two_stream_gbps'], d['tb_ok'], d['bit_err'])
PY
