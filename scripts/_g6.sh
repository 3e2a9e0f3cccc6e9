timeout 900 python -m pytest tests/test_gpu_link.py -x -q 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --no-cpu --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f Gbit/s  ms %.4f  e2e %.3f  check %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['check']))"
timeout 600 python scripts/bler_sweep.py --tbs 1024 2>&1 | head -12
