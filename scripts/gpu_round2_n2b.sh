#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('$1 n=%d value %.2f e2e %.2f (bound %.2f) llr %.2f (bound %.2f) blocking %.2f' % (d['n_gpus'], d['value'], e['value'], e['pcie_bound_value'], e['llr_input']['value'], e['llr_input']['pcie_bound_value'], e['llr_input']['blocking_value']))"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nproc
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | show "N1 chunks2"
NRLDPC_HOST_CHUNKS=2 timeout 300 $TR --master-port 29621 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu 2>/dev/null | show "N2 chunks2"
NRLDPC_HOST_CHUNKS=4 timeout 300 $TR --master-port 29622 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu 2>/dev/null | show "N2 chunks4"
OMP_NUM_THREADS=8 NRLDPC_HOST_CHUNKS=2 timeout 300 $TR --master-port 29623 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu 2>/dev/null | show "N2 chunks2 omp8"
