timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/exp_cfg4.py 2>&1 | tail -6
timeout 300 python bench.py --no-cpu --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f (single %.3f) e2e %.3f  check %s' % (d['value'], d['single_stream']['value'], d['e2e']['value'], d['check']))"
