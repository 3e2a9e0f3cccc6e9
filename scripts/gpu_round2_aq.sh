#!/bin/bash
# constant-parity layer barriers (NR_DEC_EVEN_PHASES): bounded smoke, parity, A/B against libnrldpc_ev0.so (the running phase bit)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1 || { echo "SMOKE FAILED/HUNG"; exit 1; }
timeout 420 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 200 2>&1 | tail -2
rm -f gpurun_out/aq_ab.jsonl
for rep in 1 2; do
for V in "" ev0; do
if [ -n "$V" ]; then export NRLDPC_LIB=$PWD/neoradium_b200/libnrldpc_$V.so; else unset NRLDPC_LIB; fi
timeout 200 python scripts/ab_quick.py --tbs 64 --steps 60 --tag "${V:-new}" >> gpurun_out/aq_ab.jsonl 2>>gpurun_out/aq_ab.err
timeout 200 python scripts/ab_quick.py --tbs 64 --steps 60 --rate 0.75 --tag "r075 ${V:-new}" >> gpurun_out/aq_ab.jsonl 2>>gpurun_out/aq_ab.err
done; done
python - <<'PY'
import json
for l in open('gpurun_out/aq_ab.jsonl'):
    d=json.loads(l); print(d['tag'], d['single_gbps'], d['two_stream_gbps'], d['tb_ok'], d['bit_err'])
PY
tail -2 gpurun_out/aq_ab.err
