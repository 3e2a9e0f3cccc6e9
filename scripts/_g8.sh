ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/launches_r1e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nr_decode -s 4 -c 1 -o gpurun_out/dec_r1e -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_r1e.log 2>&1
ncu --set full --clock-control none -k regex:"nr_bitstream|nr_rate_match|nr_parity_packed|nr_encode_packed|nr_rate_recover" -s 24 -c 8 -o gpurun_out/helpers_r1e -f python scripts/bench_kernels.py > gpurun_out/ncu_helpers_r1e.log 2>&1
python bench.py > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
cat gpurun_out/bench_r1e.json | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1e.json 2>gpurun_out/bench_ref_r1e.err; cat gpurun_out/bench_ref_r1e.json | cut -c1-300
ls -la gpurun_out
