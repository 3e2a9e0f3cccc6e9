python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['check'])"
