#!/bin/bash
# round-2 call K: e2e chunking, sweep batch size, final per-lifting-size table, launch list of the bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for ch in 2 3 4 6; do
NRLDPC_HOST_CHUNKS=$ch timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('chunks $ch: e2e %.3f llr %.3f value %.3f' % (d['e2e']['value'], d['e2e']['llr_input']['value'], d['value']))"
done
timeout 300 python scripts/bler_sweep.py --tbs 250000 --bg 1 --mod 16QAM --A 33672 --rate 0.6 --iters 8 --snrs=8.4,8.6,9.0 --batch-tbs 2048 2>/dev/null | head -3
timeout 300 python scripts/bler_sweep.py --tbs 250000 --bg 1 --mod 16QAM --A 33672 --rate 0.6 --iters 8 --snrs=8.4,8.6,9.0 --batch-tbs 1024 2>/dev/null | head -3
ALLZ=384,352,320,288,256,240,224,208,192,176,160,144,128,120,112,104,96,88,80,72,64,60,56,52,48,44,40,36,32,30,28,26,24,22,20,18,16,15,14,13,12,11,10,9,8,7,6,5,4,3,2
ROWS=17 ZCS=$ALLZ OUT=r02_zc_rows17.json timeout 600 python scripts/exp_zc.py > gpurun_out/k_zc17.log 2>&1; grep -c BG gpurun_out/k_zc17.log
ZCS=384,320,256,240,192,128,64,16 OUT=r02_zc_allrows.json timeout 600 python scripts/exp_zc.py > gpurun_out/k_zcall.log 2>&1; grep -c BG gpurun_out/k_zcall.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/k_ncu_bench.log 2>&1
wc -l gpurun_out/r02_launches_bench.csv
