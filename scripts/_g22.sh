timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 2>gpurun_out/bench2.err | tail -1 > gpurun_out/bench_n2.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').readline())
print('N=%d value %.3f (single %.3f)  e2e %.3f (blocking %.3f, pcie %.3f)'%(d['n_gpus'],d['value'],d['single_stream']['value'],d['e2e']['value'],d['e2e']['blocking_value'],d['e2e']['pcie_bound_value']), d['check'], d['clocks'])"
tail -3 gpurun_out/bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/bler_sweep.py --tbs 100000 --batch-tbs 512 --snrs 8.2,8.6 2>/dev/null | grep SNR
