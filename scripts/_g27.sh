timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1i.json 2>gpurun_out/bench_r1i.err; tail -3 gpurun_out/bench_r1i.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r1i.json').readline())
print('value %.3f (single %.3f)  e2e %.3f (blocking %.3f, pcie bound %.3f) f16 %.3f ok=%s'%(d['value'],d['single_stream']['value'],d['e2e']['value'],d['e2e']['blocking_value'],d['e2e']['pcie_bound_value'],d['e2e']['f16_llr_transport']['value'],d['e2e']['f16_llr_transport']['bits_ok']))
print(d['check'], d['e2e']['bits_ok'], d['clocks'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['alu_issue']['frac'])"
