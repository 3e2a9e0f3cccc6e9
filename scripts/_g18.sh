mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nr_decode -s 4 -c 1 -o gpurun_out/decode_r1h -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_f.log 2>&1
for NTB in 256 1024; do
NTB=$NTB timeout 300 python scripts/bench_kernels.py > gpurun_out/helpers_r1h_$NTB.json 2>/dev/null
done
ls -la gpurun_out | tail -8
