#!/bin/bash
# automatic early-termination start (NRLDPC_DEC_ES_AUTO): parity protocol, then throughput against from = 1 and no test
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "early_stop" 2>&1 | tail -3
rm -f gpurun_out/ab_es.jsonl
run() { timeout 300 python scripts/ab_quick.py "$@" >> gpurun_out/ab_es.jsonl 2>>gpurun_out/ab_es.err; }
for snr in 8.6 9.0 10.5; do
run --tbs 1024 --steps 5 --snr $snr --tag "noes_$snr"
run --es --es-from 1 --tbs 1024 --steps 5 --snr $snr --tag "es1_$snr"
run --es --es-from auto --tbs 1024 --steps 5 --snr $snr --tag "esauto_$snr"
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_es.jsonl'):
    d=json.loads(l); print(d['tag'], d['single_gbps'], d['two_stream_gbps'], d['tb_ok'], d['bit_err'], round(d['mean_iters'],2))
PY
tail -3 gpurun_out/ab_es.err
