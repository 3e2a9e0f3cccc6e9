python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
NRLDPC_NO_STATIC_ROWS=$v python bench.py --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nostatic=$v', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['check'])"
done
