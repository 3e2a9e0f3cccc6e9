#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_link.py -m gpu -x -q -k fused_symbol 2>&1 | grep -v "^$" | tail -4
for ch in 1 2; do for nf in 2 3 4; do
NRLDPC_HOST_CHUNKS=$ch BENCH_IN_FLIGHT=$nf timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('chunks $ch in flight $nf: e2e %.3f (bound %.2f) llr %.3f ok %s %s value %.3f' % (e['value'], e['pcie_bound_value'], e['llr_input']['value'], e['bits_ok'], e['bits_identical_to_llr_input_leg'], d['value']))"
done; done
