timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --no-cpu --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f Gbit/s  ms %.4f  e2e %.3f  check %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['check']))"
done
