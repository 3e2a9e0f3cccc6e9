#!/bin/bash
# usage: gpu_round2_multi.sh N   -- BASELINE configs[3] / [4] on N GPUs of one box (strong scaling, NCCL-reduced counters)
N=$1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "$N" = "1" ]; then TR="python"; fi
O3=gpurun_out/r02_config3_multigpu_n$N.jsonl
O4=gpurun_out/r02_config4_sweep_n$N.jsonl
rm -f $O3 $O4
timeout 600 $TR scripts/exp_cfg3_multigpu.py --out $O3 2>gpurun_out/m${N}_err.log | cut -c1-300
timeout 600 $TR scripts/exp_cfg3_multigpu.py --no-es --out $O3 2>>gpurun_out/m${N}_err.log | cut -c1-300
timeout 600 $TR scripts/exp_cfg3_multigpu.py --es-from 4 --out $O3 2>>gpurun_out/m${N}_err.log | cut -c1-300
timeout 900 $TR scripts/bler_sweep.py --tbs 250000 --bg 1 --mod 16QAM --A 33672 --rate 0.6 --iters 8 --snrs=7.0,7.4,7.8,8.0,8.2,8.4,8.6,8.8,9.0,9.4 --batch-tbs 512 --out $O4 2>>gpurun_out/m${N}_err.log | tail -12 | cut -c1-200
timeout 900 $TR scripts/bler_sweep.py --tbs 1000000 --bg 2 --mod QPSK --A 3000 --rate 0.3 --iters 8 --snrs=-3.3,-3.0,-2.7,-2.4,-2.1,-1.8,-1.5,-1.2,-0.9,-0.6 --batch-tbs 2048 --out $O4 2>>gpurun_out/m${N}_err.log | tail -12 | cut -c1-200
timeout 300 $TR scripts/h2d_ceiling.py 2>>gpurun_out/m${N}_err.log | tee gpurun_out/r02_h2d_ceiling_n$N.jsonl
timeout 300 $TR scripts/h2d_ceiling.py --affinity 2>>gpurun_out/m${N}_err.log | tee -a gpurun_out/r02_h2d_ceiling_n$N.jsonl
tail -5 gpurun_out/m${N}_err.log
