# A/B: layer barrier mode x TMA staging of the load phase
for BAR in 0 1 2; do
  for NS in 0 1; do
    echo "== BAR=$BAR NO_STAGE=$NS"
    NRLDPC_DEC_BAR=$BAR NRLDPC_NO_STAGE=$NS timeout 300 python bench.py --no-cpu --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f Gbit/s  ms %.4f  e2e %.3f  check %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['check']))"
  done
done
for BAR in 0 1 2; do
  echo "== tests BAR=$BAR"
  NRLDPC_DEC_BAR=$BAR timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
done
