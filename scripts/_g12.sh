for U in 1 2 4; do
for NTB in 256 1024; do
echo "== U=$U NTB=$NTB"
NRLDPC_CRC_UNROLL=$U NTB=$NTB python scripts/bench_kernels.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for k,v in d['stages'].items():
    if 'crc' in k or 'segment' in k: print('%-28s %8.3f ms %8.1f GB/s %5.1f%%'%(k,v['ms'],v['GBps'],100*v['frac_of_measured_hbm']))
"
done
done
