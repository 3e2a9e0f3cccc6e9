timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
timeout 300 python bench.py --no-cpu --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f (single %.3f) e2e %.3f  check %s' % (d['value'], d['single_stream']['value'], d['e2e']['value'], d['check']))"
done
