timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for NTB in 256 1024; do
echo "== NTB=$NTB ($((NTB*16)) code blocks)"
NTB=$NTB python scripts/bench_kernels.py > gpurun_out/helpers_r1f_$NTB.json 2>gpurun_out/helpers_r1f.err; python -c "
import json; d=json.load(open('gpurun_out/helpers_r1f_$NTB.json'))
for k,v in d['stages'].items(): print('%-28s %8.3f ms %8.1f GB/s %5.1f%%  %7.1f Mcb/s'%(k,v['ms'],v['GBps'],100*v['frac_of_measured_hbm'],v['Mcb_per_s']))
"
tail -3 gpurun_out/helpers_r1f.err
done
