#!/bin/bash
# round-2 call H: multi-block static kernels (parity + per-lifting-size throughput), symbols e2e test, bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -12 gpurun_out/h_pytest.log
ALLZ=384,352,320,288,256,240,224,208,192,176,160,144,128,120,112,104,96,88,80,72,64,60,56,52,48,44,40,36,32,30,28,26,24,22,20,18,16,15,14,13,12,11,10,9,8,7,6,5,4,3,2
ROWS=17 ZCS=$ALLZ OUT=r02_zc_rows17.json timeout 900 python scripts/exp_zc.py > gpurun_out/h_zc17.log 2>&1; tail -3 gpurun_out/h_zc17.log
ROWS=17 ZCS=240,208,192,128,64,32,8 NRLDPC_NO_STATIC_MB=1 OUT=r02_zc_rows17_generic.json timeout 600 python scripts/exp_zc.py > gpurun_out/h_zc17g.log 2>&1
ZCS=384,256,240,192,128,64 OUT=r02_zc_allrows.json timeout 600 python scripts/exp_zc.py > gpurun_out/h_zcall.log 2>&1; tail -12 gpurun_out/h_zcall.log
paste <(grep BG gpurun_out/h_zc17g.log | cut -c1-75) <(echo) | head -20
grep -E "Zc=(384|352|256|240|208|192|128| 64| 32|  8) " gpurun_out/h_zc17.log | cut -c1-80
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; tail -c 1500 gpurun_out/h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/h_bench.json'))
print("value", d['value'], "single", d['single_stream']['value'], "e2e", d['e2e']['value'], "llr", d['e2e']['llr_input']['value'], d['e2e']['bits_ok'], d['e2e']['bits_identical_to_llr_input_leg'], "pcie", d['e2e']['pcie_bound_value'], "frac", d['roofline']['frac'], "fp64", d['roofline']['fp64']['value'])
PY
