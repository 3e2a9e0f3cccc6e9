#!/bin/bash
cd $GRAFT_REPO_ROOT
for nf in 2 3 4; do
BENCH_IN_FLIGHT=$nf timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('in flight $nf: e2e %.3f (bound %.2f) llr %.3f (bound %.2f) f16 %.2f ok %s %s %s value %.3f' % (e['value'], e['pcie_bound_value'], e['llr_input']['value'], e['llr_input']['pcie_bound_value'], e['f16_llr_transport']['value'], e['bits_ok'], e['bits_identical_to_llr_input_leg'], e['llr_input']['bits_ok'], d['value']))"
done
