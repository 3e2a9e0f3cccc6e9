#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -8 gpurun_out/f_pytest.log
run() { timeout 300 python scripts/ab_quick.py "$@" >> gpurun_out/f_ab.jsonl 2>>gpurun_out/f_ab.err; }
rm -f gpurun_out/f_ab.jsonl
run --tag new
run --es --tbs 4096 --steps 3 --tag es65k
run --tbs 4096 --steps 3 --tag noes65k
run --es --tag es1k
run --es --snr 8.6 --tbs 1024 --steps 5 --tag es86
run --es --snr 10.5 --tbs 1024 --steps 5 --tag es105
cat gpurun_out/f_ab.jsonl | cut -c1-330
tail -3 gpurun_out/f_ab.err
