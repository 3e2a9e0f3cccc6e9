#!/bin/bash
# asynchronous Tensor-Memory loads / stores: parity, then A/B against libnrldpc_v0.so (the previous build)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 120 python scripts/dbg_mb.py 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
rm -f gpurun_out/y_ab.jsonl
for rep in 1 2; do
for V in "" v0; do
if [ -n "$V" ]; then export NRLDPC_LIB=$PWD/neoradium_b200/libnrldpc_$V.so; else unset NRLDPC_LIB; fi
timeout 200 python scripts/ab_quick.py --tbs 64 --steps 60 --tag "${V:-new}" >> gpurun_out/y_ab.jsonl 2>>gpurun_out/y_ab.err
timeout 200 python scripts/ab_quick.py --tbs 64 --steps 60 --rate 0.45 --tag "r045 ${V:-new}" >> gpurun_out/y_ab.jsonl 2>>gpurun_out/y_ab.err
timeout 200 python scripts/ab_quick.py --tbs 64 --steps 60 --rate 0.33 --tag "r033 ${V:-new}" >> gpurun_out/y_ab.jsonl 2>>gpurun_out/y_ab.err
timeout 200 python scripts/ab_quick.py --tbs 1024 --steps 5 --es --es-from 6 --tag "es6 ${V:-new}" >> gpurun_out/y_ab.jsonl 2>>gpurun_out/y_ab.err
done; done
python - <<'PY'
import json
for l in open('gpurun_out/y_ab.jsonl'):
    d=json.loads(l); print(d['tag'], d['single_gbps'], d['two_stream_gbps'], d['tb_ok'], d['bit_err'])
PY
tail -3 gpurun_out/y_ab.err
