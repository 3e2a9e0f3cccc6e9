#!/bin/bash
# round-2 call B: parity + A/B of the CRC fold / work queue / ES SpecTab, one ncu --set full capture of the Zc=384 kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
R1=$PWD/neoradium_b200/libnrldpc_r1.so
run() { timeout 300 python scripts/ab_quick.py "$@" >> gpurun_out/b_ab.jsonl 2>>gpurun_out/b_ab.err; }
run --tag new
NRLDPC_LIB=$R1 run --tag r1
run --tag new
NRLDPC_NO_DYNQ=1 run --tag nodynq
run --tbs 4096 --steps 3 --tag new65k
NRLDPC_LIB=$R1 run --tbs 4096 --steps 3 --tag r1_65k
run --es --tbs 4096 --steps 3 --tag es65k
NRLDPC_NO_DYNQ=1 run --es --tbs 4096 --steps 3 --tag es65k_nodynq
NRLDPC_LIB=$R1 run --es --tbs 4096 --steps 3 --tag es65k_r1
run --es --tag es1k
NRLDPC_NO_DYNQ=1 run --es --tag es1k_nodynq
NRLDPC_LIB=$R1 run --es --tag es1k_r1
run --es --snr 8.6 --tbs 1024 --steps 5 --tag es86
NRLDPC_LIB=$R1 run --es --snr 8.6 --tbs 1024 --steps 5 --tag es86_r1
run --rate 0.48 --tag r048
run --rate 0.40 --tag r040
NRLDPC_LIB=$R1 run --rate 0.40 --tag r040_r1
run --rate 0.75 --tag r075
NRLDPC_LIB=$R1 run --rate 0.75 --tag r075_r1
cat gpurun_out/b_ab.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nr_decode -s 6 -c 1 -f -o gpurun_out/r2b_decode python scripts/ab_quick.py --steps 3 > gpurun_out/b_ncu.log 2>&1
ls -la gpurun_out/
