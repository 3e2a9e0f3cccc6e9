#!/bin/bash
# 2-GPU sanity of the final tree: the driver's launch line for both arms
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29631 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/n2c_bench.json 2> gpurun_out/n2c_bench.err; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/n2c_bench.json').read().strip().split('\n')[-1]); e=d['e2e']; print('N2 value %.2f e2e %.2f llr %.2f ok %s %s clocks %s' % (d['value'], e['value'], e['llr_input']['value'], e['bits_ok'], e['bits_identical_to_llr_input_leg'], d['clocks']))"
timeout 400 $TR --master-port 29632 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/n2c_ref.json 2> gpurun_out/n2c_ref.err; echo "rc=$?"
tail -c 400 gpurun_out/n2c_ref.json
