#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -12 gpurun_out/e_pytest.log
timeout 600 python scripts/exp_cfg2_slot.py > gpurun_out/e_cfg2.log 2>&1; tail -5 gpurun_out/e_cfg2.log
timeout 900 python scripts/run_harq_notebook.py --transmissions 1000 --ref-transmissions 40 --out gpurun_out/r02_harq_notebook.json > gpurun_out/e_harq.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_harq_notebook.json'))
for k in d:
    if isinstance(d[k], dict): print(k, d[k].get('txBlocks_per_try'), d[k].get('rxBlocks_per_try'), d[k].get('latency_ms'))
PY
rm -f gpurun_out/r02_config3_multigpu.jsonl
timeout 600 python scripts/exp_cfg3_multigpu.py --out gpurun_out/r02_config3_multigpu.jsonl 2>gpurun_out/e_cfg3.err | cut -c1-400
timeout 600 python scripts/exp_cfg3_multigpu.py --no-es --out gpurun_out/r02_config3_multigpu.jsonl 2>>gpurun_out/e_cfg3.err | cut -c1-400
timeout 600 python scripts/exp_cfg3_multigpu.py --es-from 4 --out gpurun_out/r02_config3_multigpu.jsonl 2>>gpurun_out/e_cfg3.err | cut -c1-400
tail -3 gpurun_out/e_cfg3.err
