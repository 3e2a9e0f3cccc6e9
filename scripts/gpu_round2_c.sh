#!/bin/bash
# round-2 call C: decoder code-generation variants (A/B on the bench workload) + one bench.py line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { timeout 300 python scripts/ab_quick.py "$@" >> gpurun_out/c_ab.jsonl 2>>gpurun_out/c_ab.err; }
for rep in 1 2; do
run --tag default
for v in nofirst predmin1 fma2 fma3; do
NRLDPC_LIB=$PWD/build_ab/libnrldpc_$v.so run --tag $v
done
done
run --tbs 4096 --steps 3 --tag default65k
for v in nofirst predmin1 fma2 fma3; do
NRLDPC_LIB=$PWD/build_ab/libnrldpc_$v.so run --tbs 4096 --steps 3 --tag ${v}_65k
done
cat gpurun_out/c_ab.jsonl
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
tail -c 3000 gpurun_out/c_bench.json
