for V in 0 1 0 1; do
echo "== NRLDPC_NO_SPECZ=$V"
NRLDPC_NO_SPECZ=$V timeout 300 python bench.py --no-cpu --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f Gbit/s  ms %.4f  e2e %.3f  check %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['check']))"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
