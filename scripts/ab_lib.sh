for V in "" neoradium_b200/libnrldpc_v0.so "" neoradium_b200/libnrldpc_v0.so; do
echo "== lib=${V:-default}"
if [ -n "$V" ]; then export NRLDPC_LIB=$PWD/$V; else unset NRLDPC_LIB; fi
timeout 300 python bench.py --no-cpu --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.3f (single %.3f) e2e %.3f  check %s' % (d['value'], d['single_stream']['value'], d['e2e']['value'], d['check']))"
done
