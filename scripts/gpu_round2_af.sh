#!/bin/bash
# fused symbol input (nrldpc_decode_tb_symbols): parity, then the bench line with and without the fused demapper
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_link.py -m gpu -x -q 2>&1 | tail -5
for v in 0 1 0 1; do
NRLDPC_NO_FUSED_DEMAP=$( [ $v = 1 ] && echo 1 ) timeout 300 env $( [ $v = 1 ] && echo NRLDPC_NO_FUSED_DEMAP=1 || echo NRLDPC_DUMMY=1 ) python bench.py --steps 40 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('no_fused=$v: e2e %.3f (bound %.2f) llr %.3f ok %s %s value %.3f single %.3f' % (e['value'], e['pcie_bound_value'], e['llr_input']['value'], e['bits_ok'], e['bits_identical_to_llr_input_leg'], d['value'], d['single_stream']['value']))"
done
