mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nr_decode -s 4 -c 1 -o gpurun_out/decode_r1j -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_f.log 2>&1
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1j.json 2>gpurun_out/bench_r1j.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r1j.json').readline())
print('value %.3f (single %.3f)  e2e %.3f (blocking %.3f, pcie bound %.3f) f16 %.3f'%(d['value'],d['single_stream']['value'],d['e2e']['value'],d['e2e']['blocking_value'],d['e2e']['pcie_bound_value'],d['e2e']['f16_llr_transport']['value']))"
timeout 600 python scripts/exp_cfg4.py 2>&1 | tail -7
ls gpurun_out | tail -5
