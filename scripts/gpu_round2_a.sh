#!/bin/bash
# round-2 call A: parity of the new static schedule + A/B against the round-1 library
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
for rep in 1 2; do
for L in "" neoradium_b200/libnrldpc_r1.so; do
  if [ -n "$L" ]; then export NRLDPC_LIB=$PWD/$L; else unset NRLDPC_LIB; fi
  timeout 300 python scripts/ab_quick.py --tag rep$rep >> gpurun_out/a_ab.jsonl 2>gpurun_out/a_ab.err
  timeout 300 python scripts/ab_quick.py --tbs 1024 --steps 6 --tag rep$rep >> gpurun_out/a_ab.jsonl 2>>gpurun_out/a_ab.err
done
done
unset NRLDPC_LIB
NRLDPC_NO_STAGE=1 timeout 300 python scripts/ab_quick.py --tag nostage >> gpurun_out/a_ab.jsonl 2>>gpurun_out/a_ab.err
NRLDPC_NO_SPECZ=1 timeout 300 python scripts/ab_quick.py --tag nospecz >> gpurun_out/a_ab.jsonl 2>>gpurun_out/a_ab.err
timeout 300 python scripts/ab_quick.py --rate 0.48 --tag r048 >> gpurun_out/a_ab.jsonl 2>>gpurun_out/a_ab.err
NRLDPC_LIB=$PWD/neoradium_b200/libnrldpc_r1.so timeout 300 python scripts/ab_quick.py --rate 0.48 --tag r048 >> gpurun_out/a_ab.jsonl 2>>gpurun_out/a_ab.err
timeout 300 python scripts/ab_quick.py --es --tbs 512 --steps 10 --tag es >> gpurun_out/a_ab.jsonl 2>>gpurun_out/a_ab.err
NRLDPC_LIB=$PWD/neoradium_b200/libnrldpc_r1.so timeout 300 python scripts/ab_quick.py --es --tbs 512 --steps 10 --tag es >> gpurun_out/a_ab.jsonl 2>>gpurun_out/a_ab.err
cat gpurun_out/a_ab.jsonl
