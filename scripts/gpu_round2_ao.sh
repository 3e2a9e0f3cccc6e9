#!/bin/bash
# BG2 Zc=384, 17 rows: the current tree against the tree of commit 3891800 (build_tmp/old, its own library)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
WAVES=8 ROWS=17 ZCS=384,256 OUT=ao_new$rep.json timeout 300 python scripts/exp_zc.py > /dev/null 2>&1
(cd build_tmp/old && WAVES=8 ROWS=17 ZCS=384,256 OUT=ao_old$rep.json timeout 300 python scripts/exp_zc.py > /dev/null 2>&1; cp gpurun_out/ao_old$rep.json ../../gpurun_out/)
done
python - <<'PY'
import json
for f in ("ao_new1","ao_old1","ao_new2","ao_old2"):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, {k:round(v['g_edge_updates_per_s']) for k,v in d.items()})
PY
