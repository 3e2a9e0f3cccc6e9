mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nr_decode -s 4 -c 1 -o gpurun_out/decode -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_f.log 2>&1
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2>gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
timeout 600 python scripts/exp_cfg4.py 2>&1 | tail -7
OUT=exp_zc_allrows.json timeout 600 python scripts/exp_zc.py > gpurun_out/exp_zc_allrows.log 2>&1
ROWS=17 OUT=exp_zc_rows17.json timeout 600 python scripts/exp_zc.py > gpurun_out/exp_zc_rows17.log 2>&1
timeout 300 python scripts/exp_rate.py > gpurun_out/exp_rate.log 2>&1
tail -c 300 gpurun_out/bench.json
