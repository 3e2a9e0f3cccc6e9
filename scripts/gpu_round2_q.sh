#!/bin/bash
# R2P variant: parity first, then A/B against the previous row code (libnrldpc_v0.so = -DNR_DEC_R2P=0)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_real_callers.py -m gpu -x -q 2>&1 | tail -4
rm -f gpurun_out/q_ab.jsonl
for rep in 1 2; do
for V in "" neoradium_b200/libnrldpc_v0.so; do
if [ -n "$V" ]; then export NRLDPC_LIB=$PWD/$V; else unset NRLDPC_LIB; fi
timeout 200 python scripts/ab_quick.py --tbs 64 --steps 60 --tag "lib=${V:-default}" >> gpurun_out/q_ab.jsonl 2>>gpurun_out/q_ab.err
timeout 200 python scripts/ab_quick.py --tbs 64 --steps 60 --rate 0.33 --tag "r033 lib=${V:-default}" >> gpurun_out/q_ab.jsonl 2>>gpurun_out/q_ab.err
done; done
cut -c1-300 gpurun_out/q_ab.jsonl
tail -3 gpurun_out/q_ab.err
