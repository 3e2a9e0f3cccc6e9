#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
tail -3 gpurun_out/n2_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n2_bench.json').read().strip().split('\n')[-1])
print("N=2 value", d['value'], "single", d['single_stream']['value'], "e2e", d['e2e']['value'], d['e2e']['bits_ok'], "llr", d['e2e']['llr_input']['value'], "n_gpus", d['n_gpus'], "launches", d['gpu_launches'], "clocks", d['clocks'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | cut -c1-200
