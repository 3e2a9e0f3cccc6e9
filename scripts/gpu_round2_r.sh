#!/bin/bash
# A/B of library variants given as arguments (names of neoradium_b200/libnrldpc_NAME.so), device-resident timing
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r_ab.jsonl
for rep in 1 2; do
for V in "$@"; do
export NRLDPC_LIB=$PWD/neoradium_b200/libnrldpc_$V.so
timeout 200 python scripts/ab_quick.py --tbs 64 --steps 60 --tag "$V" >> gpurun_out/r_ab.jsonl 2>>gpurun_out/r_ab.err
done; done
python - <<'PY'
import json
for l in open('gpurun_out/r_ab.jsonl'):
    d=json.loads(l); print(d['tag'], d['single_gbps'], d['two_stream_gbps'], d['tb_ok'], d['bit_err'])
PY
tail -3 gpurun_out/r_ab.err
