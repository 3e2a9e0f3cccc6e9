#!/bin/bash
# final checks of the round: smoke, the GPU suite, the bench line, the launch list of the bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log
tail -4 gpurun_out/final_pytest.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench.json'))
print("value", round(d['value'],3), "single", round(d['single_stream']['value'],3), "e2e", round(d['e2e']['value'],3), "llr", round(d['e2e']['llr_input']['value'],3), d['e2e']['bits_ok'], d['e2e']['bits_identical_to_llr_input_leg'], "frac", round(d['roofline']['frac'],4), round(d['roofline']['two_stream_frac'],4), "fp64", round(d['roofline']['fp64']['value'],3), "cpu", d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['bits_identical_to_gpu'], d['clocks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/final_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_bench.csv | head -12 | cut -c1-200
