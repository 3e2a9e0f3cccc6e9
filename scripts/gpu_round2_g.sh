#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { timeout 300 python scripts/ab_quick.py "$@" >> gpurun_out/g_ab.jsonl 2>>gpurun_out/g_ab.err; }
rm -f gpurun_out/g_ab.jsonl
run --tbs 1024 --steps 5 --tag noes
NRLDPC_ES_CODE=1 run --tbs 1024 --steps 5 --tag noes_escode
run --es --es-from 9 --tbs 1024 --steps 5 --tag es_from9
run --es --es-from 8 --tbs 1024 --steps 5 --tag es_from8
run --es --es-from 6 --tbs 1024 --steps 5 --tag es_from6
run --es --es-from 1 --tbs 1024 --steps 5 --tag es_from1
NRLDPC_NO_DYNQ=1 run --es --es-from 1 --tbs 1024 --steps 5 --tag es_from1_nodynq
NRLDPC_NO_STAGE=1 run --es --es-from 1 --tbs 1024 --steps 5 --tag es_from1_nostage
cat gpurun_out/g_ab.jsonl | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nr_decode -s 4 -c 1 -f -o gpurun_out/r2g_decode_es python scripts/ab_quick.py --es --tbs 1024 --steps 2 > gpurun_out/g_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
