#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "early_stop or variants or row_skipping or special" 2>&1 | tail -3
run() { timeout 300 python scripts/ab_quick.py "$@" >> gpurun_out/o_ab.jsonl 2>>gpurun_out/o_ab.err; }
rm -f gpurun_out/o_ab.jsonl
run --tbs 1024 --steps 5 --tag noes
run --es --es-from 9 --tbs 1024 --steps 5 --tag es_from9
run --es --es-from 8 --tbs 1024 --steps 5 --tag es_from8
run --es --es-from 6 --tbs 1024 --steps 5 --tag es_from6
run --es --es-from 5 --tbs 1024 --steps 5 --tag es_from5
run --es --es-from 1 --tbs 1024 --steps 5 --tag es_from1
run --es --es-from 1 --snr 10.5 --tbs 1024 --steps 5 --tag es105
run --es --es-from 1 --snr 8.6 --tbs 1024 --steps 5 --tag es86
cut -c1-290 gpurun_out/o_ab.jsonl
