timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1h.json 2>gpurun_out/bench_r1h.err; tail -5 gpurun_out/bench_r1h.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r1h.json').readline())
print('value %.3f (single %.3f)  e2e %.3f (blocking %.3f, pcie bound %.3f, h2d %.3f ms)'%(d['value'],d['single_stream']['value'],d['e2e']['value'],d['e2e']['blocking_value'],d['e2e']['pcie_bound_value'],d['e2e']['h2d_only_ms_per_step']))
print(d['check'], d['e2e']['bits_ok'], d['clocks'], d['cpu_baseline'])"
for NTB in 256 1024; do
NTB=$NTB timeout 300 python scripts/bench_kernels.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for k,v in d['stages'].items():
    if 'scramble' in k: print('%-28s %8.3f ms %8.1f GB/s %5.1f%%'%(k,v['ms'],v['GBps'],100*v['frac_of_measured_hbm']))
"
done
