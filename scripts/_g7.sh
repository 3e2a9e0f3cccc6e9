timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "mixed_slot" 2>&1 | tail -5
for CH in 2 4 8 16; do
echo "== chunks $CH"
NRLDPC_HOST_CHUNKS=$CH timeout 300 python bench.py --no-cpu --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f Gbit/s  ms %.4f  e2e %.3f  check %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['bits_ok']))"
done
