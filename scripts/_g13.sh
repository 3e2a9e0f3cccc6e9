mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1g.json 2>gpurun_out/bench_r1g.err; tail -c 3000 gpurun_out/bench_r1g.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1g.json 2>/dev/null; tail -c 600 gpurun_out/bench_ref_r1g.json
for NTB in 256 1024; do
NTB=$NTB timeout 300 python scripts/bench_kernels.py > gpurun_out/helpers_r1g_$NTB.json 2>gpurun_out/helpers_r1g.err; python -c "
import json; d=json.load(open('gpurun_out/helpers_r1g_$NTB.json'))
for k,v in d['stages'].items(): print('%-28s %8.3f ms %8.1f GB/s %5.1f%%  %7.1f Mcb/s'%(k,v['ms'],v['GBps'],100*v['frac_of_measured_hbm'],v['Mcb_per_s']))
"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nr_decode -s 4 -c 1 -o gpurun_out/decode_r1g -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out
